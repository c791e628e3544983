#!/bin/bash
# one 8-GPU call: the 7B headline line at N = 8 (+ the 65B context-2048 sub-record), fused NVLink hand-off
set -u
TAG=${1:-r02u}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 100 --warmup 8 > $OUT/bench_${TAG}_n8.json 2> $OUT/bench_${TAG}_n8.err
echo "rc=$?"; python -c "
import json
d=json.loads(open('$OUT/bench_${TAG}_n8.json').read().strip().splitlines()[-1])
print('7B n8 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'parity',d.get('parity_rel_err'),'frac',d['roofline']['frac'],d['clocks'],d['config'].get('handoff'))
for k,c in (d.get('configs') or {}).items(): print(k, {kk:c.get(kk) for kk in ('value','ms_per_step','parity_rel_err','error')}, (c.get('roofline') or {}).get('frac'))" || tail -20 $OUT/bench_${TAG}_n8.err
