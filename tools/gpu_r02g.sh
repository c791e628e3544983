#!/bin/bash
# ring v2 (FP32 single), pods ring with tile-aligned assignment, Q8 ring (int8 tensor cores): parity, then A/B benches.
set -u
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
echo "=== parity (everything)"
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu_$TAG.log 2>&1; echo "rc=$?"; grep -E "rel err|worst|FAILED|Error" $OUT/pytest_gpu_$TAG.log | tail -40; tail -3 $OUT/pytest_gpu_$TAG.log
echo "=== bench FP32 A/B"
for combo in "" "LB_NO_RING=1" "LB_RING_SLOTS=5"; do
  name=$(echo "ring $combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'],d['prefill']['ms'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
env timeout 200 python tools/mega_trace.py > $OUT/trace_ring_$TAG.txt 2>&1; echo "--- trace ring"; head -14 $OUT/trace_ring_$TAG.txt
echo "=== pods"
timeout 300 python bench.py --pods 8 --steps 50 > $OUT/bench_pods8_$TAG.json 2> $OUT/bench_pods8_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench_pods8_$TAG.json'));print('[pods8] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'])" || tail -3 $OUT/bench_pods8_$TAG.err
timeout 300 python bench.py --pods 4 --steps 50 > $OUT/bench_pods4_$TAG.json 2> $OUT/bench_pods4_$TAG.err
python -c "import json;d=json.load(open('$OUT/bench_pods4_$TAG.json'));print('[pods4] value',round(d['value'],1),'frac',d['roofline']['frac'])"
echo "=== Q8"
for combo in "" "LB_NO_RING_Q8=1"; do
  name=$(echo "q8 $combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[q8 $combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'step frac',d['step_roofline']['frac'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
echo "=== 13B + full default bench line (with configs)"
timeout 300 python bench.py --model 13b --no-cpu-baseline --no-configs --steps 50 > $OUT/bench13_$TAG.json 2> $OUT/bench13_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench13_$TAG.json'));print('[13b] rc=$rc value',round(d['value'],1),'frac',d['roofline']['frac'])" || tail -3 $OUT/bench13_$TAG.err
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; echo "default bench rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_default_$TAG.json'));print('headline',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],'repeats',d['repeats'],'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'])
for k,v in (d.get('configs') or {}).items(): print(' ',k, {kk:(round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','e2e','error')}, v.get('roofline',{}).get('frac'))" || tail -5 $OUT/bench_default_$TAG.err
