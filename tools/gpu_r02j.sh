#!/bin/bash
# r02j: register-fed megakernel with the shared-memory head start of the next phase (A/B), ring with L2 look-ahead (A/B)
set -u
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
echo "=== parity of the register-fed megakernel (LB_NO_RING=1) with the head start"
for f in test_gpu_eval test_gpu_longctx test_gpu_generate; do
  LB_NO_RING=1 timeout 600 python -m pytest tests/$f.py -m gpu -q -s > $OUT/pytest_noring_${f}_$TAG.log 2>&1; echo "$f rc=$? $(tail -1 $OUT/pytest_noring_${f}_$TAG.log)"; grep -E "rel err|worst|^FAILED|Error:" $OUT/pytest_noring_${f}_$TAG.log | head -12
done
echo "=== bench FP32 A/B"
for combo in "LB_NO_RING=1" "LB_NO_RING=1 LB_MEGA_NO_PRE=1" "LB_NO_RING=1 LB_MEGA_PRE_KB=96" "LB_NO_RING=1" "LB_RING_PF_KB=0" "LB_RING_PF_KB=256" "LB_RING_PF_KB=512"; do
  name=$(echo "$combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
LB_NO_RING=1 timeout 200 python tools/mega_trace.py > $OUT/trace_mega_pre_$TAG.txt 2>&1; echo "--- trace mega + head start"; head -22 $OUT/trace_mega_pre_$TAG.txt
LB_NO_RING=1 LB_MEGA_NO_PRE=1 timeout 200 python tools/mega_trace.py > $OUT/trace_mega_nopre_$TAG.txt 2>&1; echo "--- trace mega, no head start"; head -22 $OUT/trace_mega_nopre_$TAG.txt
