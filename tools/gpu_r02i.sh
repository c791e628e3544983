#!/bin/bash
# ring v2 with the epoch fix; Q8 ring parity + phase trace; every test file in its own process.
set -u
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
echo "=== parity, one process per file"
for f in test_gpu_eval test_gpu_longctx test_gpu_q8 test_gpu_pods test_gpu_generate test_sampler_and_swap test_multi_gpu test_gpu_loader test_gpu_ops test_gpu_tc_gemm test_c_consumer; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q -s > $OUT/pytest_${f}_$TAG.log 2>&1; echo "$f rc=$? $(tail -1 $OUT/pytest_${f}_$TAG.log)"; grep -E "rel err|worst|^FAILED|Error:" $OUT/pytest_${f}_$TAG.log | head -12
done
echo "=== bench FP32 A/B"
for combo in "" "LB_NO_RING=1"; do
  name=$(echo "ring $combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'],d['prefill']['ms'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
env timeout 200 python tools/mega_trace.py > $OUT/trace_ring_$TAG.txt 2>&1; echo "--- trace ring"; head -14 $OUT/trace_ring_$TAG.txt
echo "=== Q8"
env timeout 200 python tools/mega_trace.py --q8 > $OUT/trace_q8_$TAG.txt 2>&1; echo "--- trace q8 ring"; head -14 $OUT/trace_q8_$TAG.txt
for combo in "" "LB_NO_RING_Q8=1"; do
  name=$(echo "q8 $combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --weights q8 --context 1024 --no-cpu-baseline > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[q8 $combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'step frac',d['step_roofline']['frac'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
echo "=== full default bench line (with configs)"
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; echo "default bench rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_default_$TAG.json'));print('headline',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],'repeats',d['repeats'],'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'])
for k,v in (d.get('configs') or {}).items(): print(' ',k, {kk:(round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','e2e','error')}, v.get('roofline',{}).get('frac'))" || tail -5 $OUT/bench_default_$TAG.err
echo "=== pods trace + bench"
timeout 200 python tools/pods_trace.py > $OUT/trace_pods8_$TAG.txt 2>&1; head -16 $OUT/trace_pods8_$TAG.txt
timeout 300 python bench.py --pods 8 --steps 50 > $OUT/bench_pods8_$TAG.json 2> $OUT/bench_pods8_$TAG.err; rc=$?
python -c "import json;d=json.load(open('$OUT/bench_pods8_$TAG.json'));print('[pods8] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'])" || tail -3 $OUT/bench_pods8_$TAG.err
