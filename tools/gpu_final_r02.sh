#!/bin/bash
# final single-GPU evidence of round 2: every -m gpu test, smoke(), the default bench line, launch list + ncu --set full of the decode kernel
set -u
TAG=${1:-r02v}
OUT=gpurun_out
mkdir -p $OUT
echo "=== parity, one process per file"
for f in test_gpu_eval test_gpu_longctx test_gpu_q8 test_gpu_pods test_gpu_generate test_sampler_and_swap test_multi_gpu test_gpu_loader test_gpu_ops test_gpu_tc_gemm test_c_consumer; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q > $OUT/pytest_${f}_$TAG.log 2>&1; echo "$f rc=$? $(tail -1 $OUT/pytest_${f}_$TAG.log)"; grep -E "^FAILED|Error:" $OUT/pytest_${f}_$TAG.log | head -5
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_$TAG.log
echo "=== default bench line (driver's flags)"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_default_$TAG.json'));print('headline',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],'kernel',d['roofline']['kernel'][:24],'clocks',d['clocks'],'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'],'prefill',d['prefill'])
for k,v in (d.get('configs') or {}).items(): print(' ',k, {kk:(round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','e2e','error')}, v.get('roofline',{}).get('frac'), v.get('clocks'))" || tail -5 $OUT/bench_default_$TAG.err
echo "=== launch list (plain launches, 3 decode steps after an 8-token prompt)"
LB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv python tools/profile_decode.py --steps 3 > $OUT/launches_$TAG.log 2>&1; echo "launch list rc=$?"
python - <<P
import csv,collections
rows=[r for r in csv.reader(open('$OUT/launches_$TAG.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
t=collections.Counter(); n=collections.Counter()
for r in rows[1:]:
    try: v=float(r[iv].replace(',',''))
    except: continue
    k=r[ik].split('(')[0][-60:]; t[k]+=v; n[k]+=1
tot=sum(t.values())
for k,v in t.most_common(8): print(f'{v/1e3:10.1f} us {100*v/tot:5.1f}%  x{n[k]:4d}  {k}')
P
echo "=== ncu --set full of the decode kernel (32 layers)"
LB_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_ring_kernel -s 2 -c 1 -o $OUT/prof_ring32_$TAG python tools/profile_decode.py --steps 4 > $OUT/ncu_ring32_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof_ring32_$TAG.ncu-rep --page details > $OUT/ncu_ring32_details_$TAG.txt 2>&1
ncu -i $OUT/prof_ring32_$TAG.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[-1]
for k,x in zip(h,v):
    if k in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__inst_executed.sum'): print(k,x)"
