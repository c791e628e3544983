#!/bin/bash
# r02l: row-balanced chunks for the Q8 ring (partial-tile records) and the pods ring (chunk-bounded tensor maps)
set -u
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
for f in test_gpu_q8 test_gpu_pods test_gpu_longctx test_gpu_loader; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q -s > $OUT/pytest_${f}_$TAG.log 2>&1; echo "$f rc=$? $(tail -1 $OUT/pytest_${f}_$TAG.log)"; grep -E "rel err|worst|^FAILED|Error:" $OUT/pytest_${f}_$TAG.log | head -12
done
timeout 200 python tools/mega_trace.py --q8 > $OUT/trace_q8_$TAG.txt 2>&1; echo "--- trace q8 ring"; head -14 $OUT/trace_q8_$TAG.txt
timeout 300 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8_$TAG.json 2> $OUT/bench_q8_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench_q8_$TAG.json'));print('[q8] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'step frac',d['step_roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_q8_$TAG.err
timeout 200 python tools/pods_trace.py > $OUT/trace_pods8_$TAG.txt 2>&1; echo "--- trace pods"; head -16 $OUT/trace_pods8_$TAG.txt
timeout 300 python bench.py --pods 8 --steps 50 > $OUT/bench_pods8_$TAG.json 2> $OUT/bench_pods8_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench_pods8_$TAG.json'));print('[pods8] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_pods8_$TAG.err
