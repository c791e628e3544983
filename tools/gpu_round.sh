#!/bin/bash
# One gpurun call: GPU tests, bench, ncu launch list + one full capture of the dominant kernel.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/gpu_$TAG.txt; free -g | head -2 >> $OUT/gpu_$TAG.txt; df -h /dev/shm /tmp | tail -2 >> $OUT/gpu_$TAG.txt
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu -s > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu_$TAG.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -4 $OUT/smoke_$TAG.log
echo "=== bench"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
echo "=== ncu launch list (megakernel decode path: prefill with 64 tokens, then 3 decode steps)"
LB_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 291 -c 1200 --csv --log-file $OUT/launches_mega_$TAG.csv python tools/profile_decode.py --prompt 64 --steps 3 > $OUT/ncu_launch_mega_$TAG.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_launch_mega_$TAG.log
echo "=== ncu launch list (per-op decode path, LB_NO_MEGA=1)"
LB_NO_MEGA=1 LB_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 550 -c 600 --csv --log-file $OUT/launches_perop_$TAG.csv python tools/profile_decode.py --steps 2 > $OUT/ncu_launch_perop_$TAG.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_launch_perop_$TAG.log
echo "=== ncu full (decode_mega_kernel)"
LB_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 1 -c 2 -f -o $OUT/prof_mega_$TAG python tools/profile_decode.py --steps 3 > $OUT/ncu_full_mega_$TAG.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_full_mega_$TAG.log
echo "=== ncu full (gemm_tf32x3 prefill GEMM)"
LB_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 8 -c 2 -f -o $OUT/prof_tcgemm_$TAG python tools/profile_decode.py --layers 4 --prompt 256 --steps 1 > $OUT/ncu_full_tcgemm_$TAG.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_full_tcgemm_$TAG.log
ls -la $OUT
