#!/bin/bash
# Round-2 first measurement of the int8 tensor-core Q8 GEMV (this branch only, LB_Q8_MMA=1).
set -u
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
echo "=== parity with LB_Q8_MMA=1 (Q8 tests: GPU Q8 vs the oracle on the dequantised weights)"
LB_Q8_MMA=1 timeout 600 python -m pytest tests/test_gpu_q8.py tests/test_gpu_loader.py -x -q -m gpu > $OUT/pytest_q8mma_$TAG.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest_q8mma_$TAG.log
for combo in "" "LB_Q8_MMA=1"; do
  name=$(echo "base $combo" | tr ' =' '__')
  echo "=== [$combo] bench q8"
  env $combo timeout 300 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8_${name}_$TAG.json 2> $OUT/bench_q8_${name}_$TAG.err; echo "rc=$?"; tail -2 $OUT/bench_q8_${name}_$TAG.err
  python -c "import json;d=json.load(open('$OUT/bench_q8_${name}_$TAG.json'));print('value',round(d['value'],1),{k:v['us'] for k,v in d['per_op_kernels'].items()})"
done
echo "=== Q8 megakernel (LB_Q8_MEGA=1): parity, then bench"
LB_Q8_MEGA=1 timeout 600 python -m pytest tests/test_gpu_q8.py tests/test_gpu_loader.py -x -q -m gpu > $OUT/pytest_q8mega_$TAG.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest_q8mega_$TAG.log
LB_Q8_MEGA=1 timeout 300 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8mega_$TAG.json 2> $OUT/bench_q8mega_$TAG.err; echo "rc=$?"; tail -2 $OUT/bench_q8mega_$TAG.err
python -c "import json;d=json.load(open('$OUT/bench_q8mega_$TAG.json'));print('q8 mega value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))"
