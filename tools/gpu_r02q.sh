#!/bin/bash
set -u
TAG=${1:-r02q}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_longctx.py tests/test_gpu_q8.py -m gpu -q -s > $OUT/pytest_$TAG.log 2>&1; echo "tests rc=$? $(tail -1 $OUT/pytest_$TAG.log)"; grep -E "^FAILED|Error:" $OUT/pytest_$TAG.log | head
for combo in "LB_RING_SPIN_NS=0" "LB_RING_SPIN_NS=40" "LB_RING_SPIN_NS=150" "LB_NO_RING=1"; do
  name=$(echo "$combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
timeout 200 python tools/mega_trace.py > $OUT/trace_ring_$TAG.txt 2>&1; echo "--- trace ring"; cat $OUT/trace_ring_$TAG.txt
for combo in "LB_Q8_SPIN_NS=0" "LB_Q8_SPIN_NS=40" "LB_Q8_SPIN_NS=150"; do
  name=$(echo "$combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_q8_${name}_$TAG.json 2> $OUT/bench_q8_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_q8_${name}_$TAG.json'));print('[q8 $combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'step frac',d['step_roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_q8_${name}_$TAG.err
done
timeout 300 python bench.py --model 13b --no-cpu-baseline --no-configs > $OUT/bench_13b_$TAG.json 2> $OUT/bench_13b_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench_13b_$TAG.json'));print('[13b] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_13b_$TAG.err
