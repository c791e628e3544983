#!/bin/bash
# One short gpurun call: GPU tests, smoke, default bench, megakernel phase trace, Q8 bench (cp.async ring vs register loop).
# Usage (from the repo root, on the GPU box):  bash tools/gpu_verify.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu_$TAG.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
echo "=== bench"; timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "=== mega trace"; timeout 300 python tools/mega_trace.py > $OUT/mega_trace_$TAG.txt 2>&1; echo "rc=$?"; cat $OUT/mega_trace_$TAG.txt
echo "=== bench q8 (cp.async ring)"; timeout 600 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8_$TAG.json 2> $OUT/bench_q8_$TAG.err; echo "rc=$?"; cut -c1-200 $OUT/bench_q8_$TAG.json; tail -3 $OUT/bench_q8_$TAG.err
echo "=== bench q8 (LB_Q8_SYNC=1: register-buffered loop)"; LB_Q8_SYNC=1 timeout 600 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8sync_$TAG.json 2> $OUT/bench_q8sync_$TAG.err; echo "rc=$?"; cut -c1-200 $OUT/bench_q8sync_$TAG.json
python - <<PY
import json
for n in ("bench_q8_$TAG", "bench_q8sync_$TAG"):
    try:
        d = json.load(open("$OUT/" + n + ".json")); print(n, round(d["value"], 1), {k: v["us"] for k, v in d.get("per_op_kernels", {}).items()})
    except Exception as e: print(n, "unreadable", e)
PY
