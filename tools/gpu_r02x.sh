#!/bin/bash
set -u
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_pods.py -m gpu -q -s > $OUT/pytest_pods_$TAG.log 2>&1; echo "pods tests rc=$? $(tail -1 $OUT/pytest_pods_$TAG.log)"; grep -E "rel err|^FAILED|Error:" $OUT/pytest_pods_$TAG.log | head
timeout 300 python bench.py --pods 8 --steps 50 > $OUT/bench_pods8_$TAG.json 2> $OUT/bench_pods8_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench_pods8_$TAG.json'));print('[pods8] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_pods8_$TAG.err
timeout 200 python tools/pods_trace.py > $OUT/trace_pods8_$TAG.txt 2>&1; head -16 $OUT/trace_pods8_$TAG.txt
timeout 300 python bench.py --pods 4 --steps 50 > $OUT/bench_pods4_$TAG.json 2> $OUT/bench_pods4_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench_pods4_$TAG.json'));print('[pods4] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'])" || tail -3 $OUT/bench_pods4_$TAG.err
