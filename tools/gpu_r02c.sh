#!/bin/bash
set -u
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
echo "=== legacy mma rates"; timeout 120 tools/studies/mma_rate | tee $OUT/mma_rate_$TAG.txt
echo "=== pods short run (no profiler)"; timeout 120 python tools/prof_pods.py
echo "=== ncu full on the pods megakernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega_pods -s 1 -c 1 -o $OUT/prof_pods_$TAG -f python tools/prof_pods.py > $OUT/ncu_pods_$TAG.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu_pods_$TAG.log
