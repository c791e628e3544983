#!/bin/bash
# r02o: ring kernels without divisions / clock reads on the hot path; p2p hand-off in the register-fed megakernel (default again)
set -u
TAG=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
for f in test_gpu_longctx test_gpu_eval test_gpu_q8 test_gpu_pods; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -s > $OUT/pytest_${f}_$TAG.log 2>&1; echo "$f rc=$? $(tail -1 $OUT/pytest_${f}_$TAG.log)"; grep -E "rel err|worst|^FAILED|Error:" $OUT/pytest_${f}_$TAG.log | head -16
done
for combo in "LB_RING=1" "LB_X=1"; do
  name=$(echo "$combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
LB_RING=1 timeout 200 python tools/mega_trace.py > $OUT/trace_ring_$TAG.txt 2>&1; echo "--- trace ring"; head -14 $OUT/trace_ring_$TAG.txt
timeout 300 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8_$TAG.json 2> $OUT/bench_q8_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench_q8_$TAG.json'));print('[q8] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'step frac',d['step_roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_q8_$TAG.err
timeout 200 python tools/mega_trace.py --q8 > $OUT/trace_q8_$TAG.txt 2>&1; echo "--- trace q8 ring"; head -14 $OUT/trace_q8_$TAG.txt
timeout 300 python bench.py --pods 8 --steps 50 > $OUT/bench_pods8_$TAG.json 2> $OUT/bench_pods8_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench_pods8_$TAG.json'));print('[pods8] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_pods8_$TAG.err
