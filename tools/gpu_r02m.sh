#!/bin/bash
# r02m: ring v3 (every slot consumed by all 16 warps, activation slices in registers, 12-slot ring)
set -u
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
for f in test_gpu_eval test_gpu_longctx test_gpu_generate test_sampler_and_swap; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q -s > $OUT/pytest_${f}_$TAG.log 2>&1; echo "$f rc=$? $(tail -1 $OUT/pytest_${f}_$TAG.log)"; grep -E "rel err|worst|^FAILED|Error:" $OUT/pytest_${f}_$TAG.log | head -12
done
timeout 200 python tools/mega_trace.py > $OUT/trace_ring_$TAG.txt 2>&1; echo "--- trace ring v3"; cat $OUT/trace_ring_$TAG.txt
for combo in "LB_RING=1" "LB_NO_RING=1" "LB_RING_SLOTS=8" "LB_RING=1"; do
  name=$(echo "$combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
timeout 300 python bench.py --model 13b --no-cpu-baseline --no-configs > $OUT/bench_13b_$TAG.json 2> $OUT/bench_13b_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench_13b_$TAG.json'));print('[13b] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_13b_$TAG.err
