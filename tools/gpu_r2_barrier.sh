#!/bin/bash
# Round-2 first measurement of the prepared megakernel experiments (this branch only):
#   LB_MEGA_PF=1         L2 prefetch of the next GEMV phase's first rows between barrier arrival and wait
#   LB_MEGA_WO_STATIC=1  contiguous static row split for the wo phase (=2: + software-pipelined half-batches)
# Parity first (golden logits are tolerance-based; the prefetch does not change results, the static split only the
# order in which rows are produced), then A/B of the four combinations: bench value + CTA-0 phase trace.
set -u
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
for combo in "" "LB_MEGA_PF=1" "LB_MEGA_WO_STATIC=1" "LB_MEGA_WO_STATIC=2" "LB_MEGA_PF=1 LB_MEGA_WO_STATIC=1" "LB_MEGA_PF=1 LB_MEGA_WO_STATIC=2"; do
  name=$(echo "base $combo" | tr ' =' '__')
  echo "=== [$combo] parity (eval + generate tests)"
  env $combo timeout 600 python -m pytest tests/test_gpu_eval.py tests/test_gpu_generate.py -x -q -m gpu > $OUT/pytest_${name}_$TAG.log 2>&1; echo "rc=$?"; tail -2 $OUT/pytest_${name}_$TAG.log
  echo "=== [$combo] bench"
  env $combo timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; echo "rc=$?"
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['clocks'])"
  echo "=== [$combo] phase trace"
  env $combo timeout 200 python tools/mega_trace.py > $OUT/trace_${name}_$TAG.txt 2>&1; head -15 $OUT/trace_${name}_$TAG.txt
done
