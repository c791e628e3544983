#!/bin/bash
# ring v2 (16 KB copies, row-per-warp) + pods ring with 3-D tensor-map slots; TMA copy-size study first.
set -u
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
echo "=== TMA bulk copy rate vs size"; timeout 120 tools/studies/tma_copy_rate | tee $OUT/tma_copy_rate_$TAG.txt
echo "=== parity"
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_longctx.py tests/test_gpu_pods.py tests/test_sampler_and_swap.py tests/test_c_consumer.py tests/test_gpu_generate.py -m gpu -q -s > $OUT/pytest_ring_$TAG.log 2>&1; echo "rc=$?"; grep -E "rel err|worst|Error|FAILED" $OUT/pytest_ring_$TAG.log | tail -30; tail -3 $OUT/pytest_ring_$TAG.log
echo "=== bench A/B"
for combo in "" "LB_NO_RING=1" "LB_RING_SLOTS=5"; do
  name=$(echo "ring $combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
env timeout 200 python tools/mega_trace.py > $OUT/trace_ring_$TAG.txt 2>&1; echo "--- trace ring"; head -14 $OUT/trace_ring_$TAG.txt
timeout 300 python bench.py --model 13b --no-cpu-baseline --no-configs --steps 50 > $OUT/bench13_$TAG.json 2> $OUT/bench13_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench13_$TAG.json'));print('[13b ring] rc=$rc value',round(d['value'],1),'frac',d['roofline']['frac'])" || tail -3 $OUT/bench13_$TAG.err
echo "=== pods"
for combo in "" "LB_NO_RING_PODS=1"; do
  env $combo timeout 300 python bench.py --pods 8 --steps 50 > $OUT/bench_pods8_$TAG.json 2> $OUT/bench_pods8_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_pods8_$TAG.json'));print('[pods8 $combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'])" || tail -3 $OUT/bench_pods8_$TAG.err
done
