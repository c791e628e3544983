#!/bin/bash
# Final short gpurun call of a round: GPU tests, Q8 bench, default bench (no CPU baseline leg).
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest -m gpu"; timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu_$TAG.log
echo "=== bench q8"; timeout 200 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8_$TAG.json 2> $OUT/bench_q8_$TAG.err; echo "rc=$?"
echo "=== bench"; timeout 200 python bench.py --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "rc=$?"
python - <<PY
import json
for n in ("bench_q8_$TAG", "bench_$TAG"):
    try:
        d = json.load(open("$OUT/" + n + ".json")); print(n, round(d["value"], 1), round(d["e2e"]["value"], 1), {k: v["us"] for k, v in d.get("per_op_kernels", {}).items()})
    except Exception as e: print(n, "unreadable", e)
PY
