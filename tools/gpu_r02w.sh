#!/bin/bash
set -u
TAG=${1:-r02w}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_q8.py tests/test_gpu_longctx.py -k "q8" -m gpu -q -s > $OUT/pytest_q8_$TAG.log 2>&1; echo "q8 tests rc=$? $(tail -1 $OUT/pytest_q8_$TAG.log)"; grep -E "rel err|^FAILED|Error:" $OUT/pytest_q8_$TAG.log | head
timeout 300 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8_$TAG.json 2> $OUT/bench_q8_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench_q8_$TAG.json'));print('[q8] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'step frac',d['step_roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_q8_$TAG.err
timeout 200 python tools/mega_trace.py --q8 > $OUT/trace_q8_$TAG.txt 2>&1; echo "--- trace q8 ring"; head -14 $OUT/trace_q8_$TAG.txt
