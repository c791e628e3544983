#!/bin/bash
# 8-GPU round: NCCL pipeline correctness + the BASELINE multi-GPU configs.  bash tools/gpu_multi8.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
echo "=== pytest multi-gpu"; timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > $OUT/pytest_mgpu_$TAG.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_mgpu_$TAG.log
run() {  # n model ctx steps
  local n=$1 model=$2 ctx=$3 steps=$4
  local f=$OUT/bench_${TAG}_${model}_n$n.json
  echo "=== bench --gpus $n --model $model --context $ctx"
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --model $model --context $ctx --steps $steps --no-cpu-baseline > $f 2> ${f%.json}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py --gpus $n --model $model --context $ctx --steps $steps > $f 2> ${f%.json}.err
  fi
  echo "rc=$?"; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('   value', round(d['value'],1), 'tok/s  e2e', round(d['e2e']['value'],1), ' ms/step', round(d['ms_per_step'],3), ' frac', d['roofline']['frac'])
except Exception as e: print('   parse failed', e); print(open('${f%.json}.err').read()[-1500:])
"
}
run 8 7b 512 100
run 4 7b 512 100
run 8 65b 2048 40
run 4 13b 512 60
