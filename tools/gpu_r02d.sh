#!/bin/bash
# TMA-ring megakernel: parity (eval / long context / generate / pipeline-on-one-GPU), then A/B bench and phase trace.
set -u
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
echo "=== parity with the ring megakernel (default)"
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_longctx.py tests/test_gpu_generate.py tests/test_multi_gpu.py tests/test_c_consumer.py -m gpu -q -s > $OUT/pytest_ring_$TAG.log 2>&1; echo "rc=$?"; grep -E "rel err|worst|Error" $OUT/pytest_ring_$TAG.log | tail -30; tail -3 $OUT/pytest_ring_$TAG.log
echo "=== bench A/B"
for combo in "" "LB_NO_RING=1" "LB_RING_SLOTS=6" "LB_RING_SLOTS=3"; do
  name=$(echo "ring $combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'],d['prefill'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
for combo in "" "LB_NO_RING=1"; do
  name=$(echo "ring $combo" | tr ' =' '__')
  env $combo timeout 200 python tools/mega_trace.py > $OUT/trace_${name}_$TAG.txt 2>&1; echo "--- trace [$combo]"; head -14 $OUT/trace_${name}_$TAG.txt
done
echo "=== 13B ring vs register-fed"
for combo in "" "LB_NO_RING=1"; do
  env $combo timeout 300 python bench.py --model 13b --no-cpu-baseline --no-configs --steps 50 > $OUT/bench13_$TAG.json 2> $OUT/bench13_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench13_$TAG.json'));print('[13b $combo] rc=$rc value',round(d['value'],1),'frac',d['roofline']['frac'])" || tail -3 $OUT/bench13_$TAG.err
done
echo "=== pods on the ring: parity, then bench"
timeout 600 python -m pytest tests/test_gpu_pods.py -q -s -m gpu > $OUT/pytest_pods_$TAG.log 2>&1; echo "rc=$?"; grep -E "rel err|worst|Error|error" $OUT/pytest_pods_$TAG.log | tail -20; tail -3 $OUT/pytest_pods_$TAG.log
for combo in "" "LB_NO_RING_PODS=1"; do
  env $combo timeout 300 python bench.py --pods 8 --steps 50 > $OUT/bench_pods8_$TAG.json 2> $OUT/bench_pods8_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_pods8_$TAG.json'));print('[pods8 $combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'])" || tail -3 $OUT/bench_pods8_$TAG.err
done
timeout 300 python bench.py --pods 4 --steps 50 > $OUT/bench_pods4_$TAG.json 2> $OUT/bench_pods4_$TAG.err
python -c "import json;d=json.load(open('$OUT/bench_pods4_$TAG.json'));print('[pods4] value',round(d['value'],1),'frac',d['roofline']['frac'])"
