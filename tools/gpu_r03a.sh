#!/bin/bash
set -u
TAG=${1:-r03a}
OUT=gpurun_out
mkdir -p $OUT
for combo in "LB_RING_SPIN_NS=0" "LB_RING_SPIN_NS=50" "LB_RING_SPIN_NS=0" "LB_RING_SPIN_NS=50" "LB_RING_SPIN_NS=200"; do
  name=$(echo "$combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
timeout 300 python -m pytest tests/test_gpu_longctx.py -k "ring" -m gpu -q > $OUT/pytest_ring_$TAG.log 2>&1; echo "ring tests (default spin) rc=$? $(tail -1 $OUT/pytest_ring_$TAG.log)"
