#!/bin/bash
set -u
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks_$TAG.csv &
SMI=$!
echo "--- probe ring"; LB_RING=1 timeout 300 python tools/ring_sync_probe.py 2>&1 | tail -4
echo "--- probe mega"; timeout 300 python tools/ring_sync_probe.py 2>&1 | tail -4
for combo in "LB_RING=1" "LB_X=1" "LB_RING=1" "LB_X=1"; do
  name=$(echo "$combo" | tr ' =' '__')
  env $combo timeout 300 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; rc=$?
  python -c "import json;d=json.load(open('$OUT/bench_${name}_$TAG.json'));print('[$combo] rc=$rc value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',d['roofline']['frac'],d['clocks'])" || tail -3 $OUT/bench_${name}_$TAG.err
done
kill $SMI
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/clocks_r02p.csv'))][1:]
print('clock samples',len(rows),'min/max MHz',min(int(r[0].split()[0]) for r in rows),max(int(r[0].split()[0]) for r in rows),'max W',max(float(r[1].split()[0]) for r in rows))
P
