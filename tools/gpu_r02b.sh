#!/bin/bash
# Round 2, second GPU call: full GPU test suite (long-context parity + pods megakernel), pods bench A/B.
set -u
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pods tests first (new kernel)"
timeout 600 python -m pytest tests/test_gpu_pods.py -q -s -m gpu > $OUT/pytest_pods_$TAG.log 2>&1; echo "rc=$?"; grep -E "rel err|worst|Error|error" $OUT/pytest_pods_$TAG.log | tail -20; tail -3 $OUT/pytest_pods_$TAG.log
echo "=== pytest -m gpu (everything else)"
timeout 1200 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_pods.py > $OUT/pytest_gpu_$TAG.log 2>&1; echo "rc=$?"; grep -E "rel err|worst" $OUT/pytest_gpu_$TAG.log | tail -40; tail -5 $OUT/pytest_gpu_$TAG.log
echo "=== pods bench"
for B in 8 4; do
  timeout 300 python bench.py --pods $B --steps 50 > $OUT/bench_pods${B}_$TAG.json 2> $OUT/bench_pods${B}_$TAG.err; echo "rc=$?"; tail -2 $OUT/bench_pods${B}_$TAG.err
  python -c "import json;d=json.load(open('$OUT/bench_pods${B}_$TAG.json'));print('pods$B value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])"
done
LB_NO_MEGA_PODS=1 timeout 300 python bench.py --pods 8 --steps 50 > $OUT/bench_pods8_perop_$TAG.json 2> /dev/null
python -c "import json;d=json.load(open('$OUT/bench_pods8_perop_$TAG.json'));print('pods8 per-op value',round(d['value'],1))"
