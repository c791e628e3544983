#!/bin/bash
# Multi-GPU round: NCCL pipeline test + bench at N GPUs.  Usage: bash tools/gpu_multi.sh N [tag]
set -u
N=${1:-2}
TAG=${2:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/mgpu_${TAG}_n$N.txt 2>&1
nvidia-smi topo -m >> $OUT/mgpu_${TAG}_n$N.txt 2>&1
echo "=== pytest multi-gpu"; timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -s > $OUT/pytest_mgpu_${TAG}_n$N.log 2>&1; echo "rc=$?"; tail -12 $OUT/pytest_mgpu_${TAG}_n$N.log
EXTRA=${EXTRA:-}
for n in $(echo $N | tr ',' ' '); do
  echo "=== bench --gpus $n"
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --no-cpu-baseline > $OUT/bench_${TAG}_n$n.json 2> $OUT/bench_${TAG}_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n $EXTRA > $OUT/bench_${TAG}_n$n.json 2> $OUT/bench_${TAG}_n$n.err
  fi
  echo "rc=$?"; tail -2 $OUT/bench_${TAG}_n$n.json; tail -5 $OUT/bench_${TAG}_n$n.err
done
