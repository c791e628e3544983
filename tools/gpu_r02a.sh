#!/bin/bash
# Round 2, first GPU call: full GPU test suite (incl. the new long-context parity tests), then the first
# measurement of the prepared experiments (barrier prefetch / static wo; int8-mma Q8 GEMV and Q8 megakernel), and the
# pods-8 baseline.
set -u
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu_$TAG.log 2>&1; echo "rc=$?"; grep -E "rel err|worst" $OUT/pytest_gpu_$TAG.log | tail -30; tail -3 $OUT/pytest_gpu_$TAG.log
echo "=== megakernel experiments"
for combo in "" "LB_MEGA_PF=1" "LB_MEGA_WO_STATIC=1" "LB_MEGA_WO_STATIC=2" "LB_MEGA_PF=1 LB_MEGA_WO_STATIC=1" "LB_MEGA_PF=1 LB_MEGA_WO_STATIC=2"; do
  name=$(echo "base $combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['clocks'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
for combo in "" "LB_MEGA_PF=1" "LB_MEGA_PF=1 LB_MEGA_WO_STATIC=1"; do
  name=$(echo "base $combo" | tr ' =' '__')
  env $combo timeout 200 python tools/mega_trace.py > $OUT/trace_${name}_$TAG.txt 2>&1; echo "--- trace [$combo]"; head -12 $OUT/trace_${name}_$TAG.txt
done
echo "=== Q8 experiments"
LB_Q8_MMA=1 timeout 600 python -m pytest tests/test_gpu_q8.py tests/test_gpu_loader.py -x -q -m gpu > $OUT/pytest_q8mma_$TAG.log 2>&1; echo "q8mma parity rc=$?"; tail -3 $OUT/pytest_q8mma_$TAG.log
LB_Q8_MEGA=1 timeout 600 python -m pytest tests/test_gpu_q8.py tests/test_gpu_loader.py -x -q -m gpu > $OUT/pytest_q8mega_$TAG.log 2>&1; echo "q8mega parity rc=$?"; tail -3 $OUT/pytest_q8mega_$TAG.log
for combo in "" "LB_Q8_MMA=1" "LB_Q8_MEGA=1"; do
  name=$(echo "base $combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8_${name}_$TAG.json 2> $OUT/bench_q8_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_q8_${name}_$TAG.json'));print('[q8 $combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),{k:v['us'] for k,v in d['per_op_kernels'].items()})" || tail -3 $OUT/bench_q8_${name}_$TAG.err
done
echo "=== pods 8 baseline"
timeout 300 python bench.py --pods 8 --steps 50 > $OUT/bench_pods8_$TAG.json 2> $OUT/bench_pods8_$TAG.err; echo "rc=$?"
python -c "import json;d=json.load(open('$OUT/bench_pods8_$TAG.json'));print('pods8 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['roofline']['frac'])"
