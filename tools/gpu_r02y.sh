#!/bin/bash
set -u
TAG=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
for combo in "LB_RING_WO_L2=0" "LB_RING_WO_L2=1" "LB_RING_WO_L2=0" "LB_RING_WO_L2=1"; do
  name=$(echo "$combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
LB_RING_WO_L2=1 timeout 200 python tools/mega_trace.py > $OUT/trace_ring_wo_l2_$TAG.txt 2>&1; echo "--- trace ring, wo L2 prefetch"; head -14 $OUT/trace_ring_wo_l2_$TAG.txt
LB_RING_WO_L2=1 timeout 300 python -m pytest tests/test_gpu_longctx.py -k "ring" -m gpu -q > $OUT/pytest_ring_$TAG.log 2>&1; echo "ring tests rc=$? $(tail -1 $OUT/pytest_ring_$TAG.log)"
