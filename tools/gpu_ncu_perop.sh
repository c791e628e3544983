#!/bin/bash
# ncu --set full captures of the per-op decode kernels (Q8 weights) and of the Q8 tcgen05 prefill GEMM.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
echo "=== ncu full: one decode layer of the per-op Q8 path (rmsnorm, qkv GEMV, rope+KV store, attention, wo, rmsnorm, w1/w3 SwiGLU, w2)"
LB_NO_GRAPH=1 timeout 500 ncu --set full --clock-control none --import-source on \
  -k regex:'gemv_q8_db|attention_decode|rms_norm|rope_qk_store' -s 68 -c 8 -f -o $OUT/prof_perop_q8_$TAG \
  python tools/profile_decode.py --weights q8 --layers 4 --prompt 1 --steps 3 --ctx 1024 > $OUT/ncu_perop_q8_$TAG.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_perop_q8_$TAG.log
echo "=== ncu full: Q8 tcgen05 prefill GEMM (dequant fused in the smem stage)"
LB_NO_GRAPH=1 timeout 500 ncu --set full --clock-control none --import-source on \
  -k regex:gemm_tf32x3 -s 5 -c 2 -f -o $OUT/prof_tcgemm_q8_$TAG \
  python tools/profile_decode.py --weights q8 --layers 2 --prompt 256 --steps 1 --ctx 512 > $OUT/ncu_tcgemm_q8_$TAG.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_tcgemm_q8_$TAG.log
ls -la $OUT | tail -8
