#!/bin/bash
# Short gpurun call: GPU tests + Q8 decode bench, double-buffered GEMV vs the single-buffered loop (LB_Q8_SYNC=1).
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu_$TAG.log
echo "=== bench q8 (double-buffered)"; timeout 600 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8_$TAG.json 2> $OUT/bench_q8_$TAG.err; echo "rc=$?"; tail -3 $OUT/bench_q8_$TAG.err
echo "=== bench q8 (LB_Q8_SYNC=1)"; LB_Q8_SYNC=1 timeout 600 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8sync_$TAG.json 2> $OUT/bench_q8sync_$TAG.err; echo "rc=$?"
python - <<PY
import json
for n in ("bench_q8_$TAG", "bench_q8sync_$TAG"):
    try:
        d = json.load(open("$OUT/" + n + ".json")); print(n, round(d["value"], 1), {k: v["us"] for k, v in d.get("per_op_kernels", {}).items()})
    except Exception as e: print(n, "unreadable", e)
PY
