#!/bin/bash
# ncu --set full of the three ring megakernels (8-layer 7B-shaped models keep a launch short), summaries into gpurun_out/
set -u
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
M='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|gpu__dram_throughput|sm__warps_active|smsp__issue_active|l1tex__data_pipe_lsu_wavefronts_mem_shared|smsp__average_warp.*issue_stalled|launch__registers|smsp__inst_executed.sum |sm__inst_executed_pipe|smsp__warp_issue_stalled.*per_warp_active|lts__t_sectors_srcunit_tex_op_read|sm__pipe_tensor|l1tex__data_bank_conflicts_pipe_lsu_mem_shared'
LB_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_ring_q8 -s 2 -c 1 -o $OUT/prof_q8ring_$TAG python tools/profile_decode.py --layers 8 --weights q8 --steps 4 > $OUT/ncu_q8ring_$TAG.log 2>&1; echo "ncu q8 rc=$?"
ncu -i $OUT/prof_q8ring_$TAG.ncu-rep --page details > $OUT/ncu_q8ring_details_$TAG.txt 2>&1
ncu -i $OUT/prof_q8ring_$TAG.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,re
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[-1]
for k,x in zip(h,v):
    if re.search(r'$M',k): print(k,x)" | head -80
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_ring_pods -s 1 -c 1 -o $OUT/prof_podsring_$TAG python tools/prof_pods.py --layers 8 --steps 3 > $OUT/ncu_podsring_$TAG.log 2>&1; echo "ncu pods rc=$?"
ncu -i $OUT/prof_podsring_$TAG.ncu-rep --page details > $OUT/ncu_podsring_details_$TAG.txt 2>&1
LB_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_ring_kernel -s 2 -c 1 -o $OUT/prof_ring_$TAG python tools/profile_decode.py --layers 8 --steps 4 > $OUT/ncu_ring_$TAG.log 2>&1; echo "ncu ring rc=$?"
ncu -i $OUT/prof_ring_$TAG.ncu-rep --page details > $OUT/ncu_ring_details_$TAG.txt 2>&1
ls -la $OUT/*.ncu-rep
