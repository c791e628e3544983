// kernels_q8.cu — block-quantised INT8 weights x FP32 activations (BASELINE config 3).
//
// The reference has no quantised path (only the enum space ml.go:85-94 and the block constants
// QK = 32, ml.go:24,123-124; INT8 is an unchecked roadmap item, README.md:45), so the format is
// defined here consistently with those constants:
//   block = 32 consecutive weights of one row (along K);  d = max|w| / 127  (FP32);
//   q_i = rint(w_i / d) clamped to [-127, 127] (FP32 divide, round-half-even); d == 0 -> q = 0.
//   36 bytes per 32 weights.
// Physical layout in HBM ("4-row interleaved", chosen for the decode GEMV): rows are grouped by 4;
//   q plane: group g, k4 = k/4  ->  16 bytes at ((g*(K/4) + k4)*16): byte (r%4)*4 + (k%4)
//   d plane: group g, kb = k/32 ->  4 floats at ((g*(K/32) + kb)*4): float (r%4)
// so one 128-bit load gives a lane 4 rows x 4 consecutive weights, which all meet the SAME float4 of
// the activation vector (one fully coalesced 512-byte activation request per warp serves 16 weights
// per lane), and one more 128-bit load gives the 4 rows' block scales.  A sub-matrix that starts at a
// row multiple of 4 starts at q + r0*K, d + r0*K/32 like in a plain row-major layout.
// Parity target: the FP32 path on the dequantised weights f32(d * q_i); the GEMV computes
// d * sum_4(q_i * x_i) per 4 weights (reassociated scale), the GEMM f32(d*q_i) exactly.
#include "common.cuh"
#include "kernels.cuh"

#include <cstdlib>
#include <type_traits>

namespace lb {
namespace k {

__device__ __forceinline__ size_t q8_qoff(uint32_t r, uint32_t kcol, uint32_t K) {
    return ((size_t)(r >> 2) * (K >> 2) + (kcol >> 2)) * 16 + (r & 3) * 4 + (kcol & 3);
}
__device__ __forceinline__ size_t q8_doff(uint32_t r, uint32_t kcol, uint32_t K) {
    return ((size_t)(r >> 2) * (K >> 5) + (kcol >> 5)) * 4 + (r & 3);
}

// one thread per (row group, 32-block): 4 rows x 32 weights
__global__ void quantize_q8_kernel(const float *__restrict__ W, int8_t *__restrict__ q, float *__restrict__ d, uint32_t rows,
                                   uint32_t K) {
    const uint32_t KB = K >> 5;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)(rows >> 2) * KB) return;
    const uint32_t g = (uint32_t)(idx / KB), kb = (uint32_t)(idx % KB);
    float dd[4];
    uint32_t packed[4][8];  // [row][k4] 4 int8 each
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const float4 *src = reinterpret_cast<const float4 *>(W + (size_t)(g * 4 + rr) * K + kb * 32);
        float v[32];
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float4 t = src[i];
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
        }
        dd[rr] = __fdiv_rn(amax, 127.0f);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int r = dd[rr] > 0.f ? __float2int_rn(__fdiv_rn(v[4 * i + j], dd[rr])) : 0;
                r = max(-127, min(127, r));
                w |= ((uint32_t)(r & 0xFF)) << (8 * j);
            }
            packed[rr][i] = w;
        }
    }
    uint4 *qdst = reinterpret_cast<uint4 *>(q + ((size_t)g * (K >> 2) + kb * 8) * 16);
#pragma unroll
    for (int i = 0; i < 8; i++) qdst[i] = make_uint4(packed[0][i], packed[1][i], packed[2][i], packed[3][i]);
    *reinterpret_cast<float4 *>(d + ((size_t)g * KB + kb) * 4) = make_float4(dd[0], dd[1], dd[2], dd[3]);
}
void quantize_q8(const float *W, int8_t *q, float *d, uint32_t rows, uint32_t K, cudaStream_t st) {
    LB_CHECK(K % 32 == 0 && rows % 4 == 0, "quantize_q8: K must be a multiple of 32 and the row count a multiple of 4");
    const size_t n = (size_t)(rows / 4) * (K / 32);
    if (!n) return;
    quantize_q8_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(W, q, d, rows, K);
    LB_LAUNCH_CHECK();
}

__global__ void dequantize_q8_kernel(const int8_t *__restrict__ q, const float *__restrict__ d, float *__restrict__ out,
                                     uint32_t rows, uint32_t K) {
    const size_t n = (size_t)rows * K, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t r = (uint32_t)(i / K), kcol = (uint32_t)(i % K);
        out[i] = __fmul_rn(d[q8_doff(r, kcol, K)], (float)q[q8_qoff(r, kcol, K)]);
    }
}
void dequantize_q8(const int8_t *q, const float *d, float *out, uint32_t rows, uint32_t K, cudaStream_t st) {
    const size_t n = (size_t)rows * K;
    if (!n) return;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    dequantize_q8_kernel<<<(unsigned)blocks, 256, 0, st>>>(q, d, out, rows, K);
    LB_LAUNCH_CHECK();
}

// int8 -> float without the slow I2F pipe: bytes are biased to unsigned (xor 0x80), spliced into the
// mantissa of 2^23 with PRMT, and the bias (2^23 + 128) is subtracted — exact for every int8.
__device__ __forceinline__ void unpack4(uint32_t w, float f[4]) {
    const uint32_t u = w ^ 0x80808080u;
    f[0] = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7650)) - 8388736.0f;
    f[1] = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7651)) - 8388736.0f;
    f[2] = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7652)) - 8388736.0f;
    f[3] = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7653)) - 8388736.0f;
}
__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Decode GEMV over the interleaved layout.  A warp owns one row group (4 rows) of W1 (and of W3 for
// SwiGLU), or — KSPLIT = 4, for matrices with few rows — a quarter of its K range, the four warps of a
// block then combining through shared memory.  Per step a lane issues one 128-bit weight load (4 rows x
// 4 weights), one 128-bit scale load and one 128-bit activation load per column; Q8_UNROLL steps in flight.
constexpr int Q8_WARPS = 4;
constexpr int Q8_UNROLL = 4;

template <int NC, bool SWIGLU, int KSPLIT>
__global__ void __launch_bounds__(Q8_WARPS * 32)
gemv_q8_kernel(const int8_t *__restrict__ Q1, const float *__restrict__ D1, const int8_t *__restrict__ Q3,
               const float *__restrict__ D3, uint32_t M, uint32_t K, const float *__restrict__ x, uint32_t ldx,
               float *__restrict__ y, uint32_t ldy, const float *__restrict__ res) {
    constexpr int NM = SWIGLU ? 2 : 1;
    __shared__ float part[KSPLIT > 1 ? Q8_WARPS : 1][NM][4][NC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g = KSPLIT > 1 ? blockIdx.x : blockIdx.x * Q8_WARPS + warp;  // row group
    const bool active = g < (M >> 2);
    const uint32_t K4 = K >> 2, KB = K >> 5;
    // this warp's k4 range
    const uint32_t per = KSPLIT > 1 ? ((K4 / KSPLIT + 31) & ~31u) : K4;
    const uint32_t k4_begin = KSPLIT > 1 ? min(warp * per, K4) : 0, k4_end = KSPLIT > 1 ? min(k4_begin + per, K4) : K4;
    const uint4 *q1 = reinterpret_cast<const uint4 *>(Q1) + (size_t)(active ? g : 0) * K4;
    const float4 *d1 = reinterpret_cast<const float4 *>(D1) + (size_t)(active ? g : 0) * KB;
    const uint4 *q3 = SWIGLU ? reinterpret_cast<const uint4 *>(Q3) + (size_t)(active ? g : 0) * K4 : nullptr;
    const float4 *d3 = SWIGLU ? reinterpret_cast<const float4 *>(D3) + (size_t)(active ? g : 0) * KB : nullptr;
    float acc[NM][4][NC];
#pragma unroll
    for (int m = 0; m < NM; m++)
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < NC; c++) acc[m][r][c] = 0.f;

    uint4 w[Q8_UNROLL][NM];
    float4 s[Q8_UNROLL][NM];
    auto load_batch = [&](uint32_t kk) {
#pragma unroll
        for (int u = 0; u < Q8_UNROLL; u++) {
            const uint32_t k4 = kk + u * 32;
            const bool ok = active && k4 < k4_end;
            w[u][0] = ok ? ld_stream_u4(q1 + k4) : make_uint4(0x80808080u ^ 0x80808080u, 0, 0, 0);
            s[u][0] = ok ? __ldg(d1 + (k4 >> 3)) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (SWIGLU) {
                w[u][NM - 1] = ok ? ld_stream_u4(q3 + k4) : make_uint4(0, 0, 0, 0);
                s[u][NM - 1] = ok ? __ldg(d3 + (k4 >> 3)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    pdl_launch_dependents();
    load_batch(k4_begin + lane);  // read-only weights first, then wait for the predecessor grid (PDL)
    pdl_wait();
    for (uint32_t kk = k4_begin + lane; kk < k4_end;) {
#pragma unroll
        for (int u = 0; u < Q8_UNROLL; u++) {
            const uint32_t k4 = kk + u * 32;
            if (k4 < k4_end) {
                float4 xv[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) xv[c] = __ldg(reinterpret_cast<const float4 *>(x + (size_t)c * ldx) + k4);
#pragma unroll
                for (int m = 0; m < NM; m++) {
                    const uint32_t wr[4] = {w[u][m].x, w[u][m].y, w[u][m].z, w[u][m].w};
                    const float sr[4] = {s[u][m].x, s[u][m].y, s[u][m].z, s[u][m].w};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        float f[4];
                        unpack4(wr[r], f);
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            float t = f[0] * xv[c].x;
                            t = fmaf(f[1], xv[c].y, t); t = fmaf(f[2], xv[c].z, t); t = fmaf(f[3], xv[c].w, t);
                            acc[m][r][c] = fmaf(sr[r], t, acc[m][r][c]);
                        }
                    }
                }
            }
        }
        kk += 32 * Q8_UNROLL;
        if (kk < k4_end) load_batch(kk);
    }
#pragma unroll
    for (int m = 0; m < NM; m++)
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < NC; c++) acc[m][r][c] = warp_sum(acc[m][r][c]);
    if (KSPLIT > 1) {
        if (lane == 0) {
#pragma unroll
            for (int m = 0; m < NM; m++)
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < NC; c++) part[warp][m][r][c] = acc[m][r][c];
        }
        __syncthreads();
        if (warp != 0) return;
#pragma unroll
        for (int m = 0; m < NM; m++)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    float t = part[0][m][r][c];
                    for (int wv = 1; wv < Q8_WARPS; wv++) t += part[wv][m][r][c];
                    acc[m][r][c] = t;
                }
    }
    if (lane == 0 && active) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t row = g * 4 + r;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                float v = SWIGLU ? __fmul_rn(silu_ref(acc[0][r][c]), acc[NM - 1][r][c]) : acc[0][r][c];
                if (!SWIGLU && res) v = __fadd_rn(v, res[(size_t)c * ldy + row]);
                y[(size_t)c * ldy + row] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// EXPERIMENT (LB_Q8_MMA=1, unmeasured): the Q8 GEMV on the int8 tensor cores.  The activation vector is turned once
// per GEMV into 4 balanced base-128 digits per element relative to its Q8 block's power-of-two scale
// (q8_digits_kernel; x = 2^e * sum_j dig_j 128^-(j+1), exact to 2^-28 of the block maximum), stored as ready-made B
// fragments of `mma.sync.m16n8k32.s8` (digit j = column j; columns 4..7 are zero).  The weights go from HBM into
// the A fragment AS THEY ARE: in the 4-row interleaved planes one 32-bit word is 4 consecutive k of one row, which
// is exactly one A register.  Per (16 rows x 32 k) tile: 4 LDG.32 + 1 MMA + 4 I2F + ~8 FP32 ops instead of
// 3.5 instructions per weight; the s32 dot products are exact.  Index math checked on the CPU, lane by lane, in
// tools/studies/q8_mma_layout_emulation.py; numerics in tools/studies/q8_int8_digits.py.
__device__ __forceinline__ uint32_t ld_stream_u32(const void *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// one warp per Q8 block of the activation vector: lane k holds x[32 b + k]
__global__ void __launch_bounds__(128) q8_digits_kernel(const float *__restrict__ x, uint32_t K, uint2 *__restrict__ bfrag,
                                                        float *__restrict__ xs) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= (K >> 5)) return;
    const float v = x[(size_t)b * 32 + lane];
    const float mx = warp_max(fabsf(v));
    uint32_t pack = 0;  // digit j in byte j
    float scale = 0.f;
    if (mx >= 1e-30f && mx <= 1e30f) {  // outside: the block contributes nothing (or is not finite)
        const int e = ilogbf(mx) + 2;   // |v| / 2^e < 0.5
        scale = ldexpf(1.0f, e);
        float r = v * ldexpf(1.0f, -e);  // exact
#pragma unroll
        for (int j = 0; j < 4; j++) {
            r *= 128.0f;                 // exact
            const float dj = rintf(r);   // |dj| <= 64
            r -= dj;                     // exact
            pack |= ((uint32_t)(int)dj & 0xffu) << (8 * j);
        }
    }
    // B fragment of lane L = (gid, tig): column gid (< 4: digit gid), rows tig*4..+3 (b0) and 16+tig*4..+3 (b1)
    const int gid = lane >> 2, tig = lane & 3;
    uint32_t b0 = 0, b1 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t lo = __shfl_sync(0xffffffffu, pack, tig * 4 + i);
        const uint32_t hi = __shfl_sync(0xffffffffu, pack, 16 + tig * 4 + i);
        b0 |= ((lo >> (8 * (gid & 3))) & 0xffu) << (8 * i);
        b1 |= ((hi >> (8 * (gid & 3))) & 0xffu) << (8 * i);
    }
    if (gid >= 4) { b0 = 0; b1 = 0; }
    bfrag[(size_t)b * 32 + lane] = make_uint2(b0, b1);
    if (lane == 0) xs[b] = scale;
}

constexpr int Q8M_U = 4;  // blocks (16 rows x 32 k tiles) in flight per warp

template <bool SWIGLU, int KS>  // KS warps split the K blocks of one 16-row tile (x 2 matrices for SwiGLU)
__global__ void __launch_bounds__(32 * KS * (SWIGLU ? 2 : 1))
gemv_q8_mma_kernel(const int8_t *__restrict__ Q1, const float *__restrict__ D1, const int8_t *__restrict__ Q3,
                   const float *__restrict__ D3, uint32_t M, uint32_t K, const uint2 *__restrict__ bfrag,
                   const float *__restrict__ xs, float *__restrict__ y, const float *__restrict__ res) {
    constexpr int NM = SWIGLU ? 2 : 1;
    __shared__ float part[NM][KS][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const int mtx = SWIGLU ? warp / KS : 0, kw = warp % KS;
    const uint32_t R0 = blockIdx.x * 16, NB = K >> 5, K4 = K >> 2;
    const uint32_t per = (NB + KS - 1) / KS, b_begin = min((uint32_t)kw * per, NB), b_end = min(b_begin + per, NB);
    const int8_t *Q = mtx ? Q3 : Q1;
    const float *D = mtx ? D3 : D1;
    const uint32_t r_lo = R0 + gid, r_hi = r_lo + 8;
    // 32-bit word of (row r, k4) in the q plane: ((r/4) * K4 + k4) * 4 + r%4;  this lane's k4 = 8 b + tig (+4)
    const uint32_t *qa = reinterpret_cast<const uint32_t *>(Q) + ((size_t)(r_lo >> 2) * K4 + tig) * 4 + (r_lo & 3);
    const uint32_t *qb = reinterpret_cast<const uint32_t *>(Q) + ((size_t)(r_hi >> 2) * K4 + tig) * 4 + (r_hi & 3);
    const float *da = D + (size_t)(r_lo >> 2) * NB * 4 + (r_lo & 3);
    const float *db = D + (size_t)(r_hi >> 2) * NB * 4 + (r_hi & 3);
    // digit weights of this lane's two columns (tig 0: digits 0,1; tig 1: digits 2,3; tig 2,3: zero columns)
    const float w0 = tig == 0 ? 0x1p-7f : 0x1p-21f, w1 = w0 * 0x1p-7f;
    float acc_lo = 0.f, acc_hi = 0.f;
    pdl_launch_dependents();
    bool waited = false;
    for (uint32_t bb = b_begin; bb < b_end; bb += Q8M_U) {
        uint32_t a[Q8M_U][4];
        float s_lo[Q8M_U], s_hi[Q8M_U], xsc[Q8M_U];
        uint2 bf[Q8M_U];
#pragma unroll
        for (int u = 0; u < Q8M_U; u++) {
            const uint32_t b = bb + u;
            const bool ok = b < b_end;
            const size_t w = (size_t)b * 32;  // 8 units of 4 words per block
            a[u][0] = ok ? ld_stream_u32(qa + w) : 0u;
            a[u][1] = ok ? ld_stream_u32(qb + w) : 0u;
            a[u][2] = ok ? ld_stream_u32(qa + w + 16) : 0u;
            a[u][3] = ok ? ld_stream_u32(qb + w + 16) : 0u;
            s_lo[u] = ok ? __ldg(da + (size_t)b * 4) : 0.f;
            s_hi[u] = ok ? __ldg(db + (size_t)b * 4) : 0.f;
        }
        if (!waited) { pdl_wait(); waited = true; }  // weights first, then the digits written by the predecessor grid
#pragma unroll
        for (int u = 0; u < Q8M_U; u++) {
            const uint32_t b = bb + u;
            const bool ok = b < b_end;
            bf[u] = ok ? __ldg(bfrag + (size_t)b * 32 + lane) : make_uint2(0u, 0u);
            xsc[u] = ok ? __ldg(xs + b) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < Q8M_U; u++) {
            int c0, c1, c2, c3;
            asm volatile(
                "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3)
                : "r"(a[u][0]), "r"(a[u][1]), "r"(a[u][2]), "r"(a[u][3]), "r"(bf[u].x), "r"(bf[u].y), "r"(0));
            const float v_lo = fmaf((float)c0, w0, (float)c1 * w1), v_hi = fmaf((float)c2, w0, (float)c3 * w1);
            acc_lo = fmaf(s_lo[u] * xsc[u], v_lo, acc_lo);
            acc_hi = fmaf(s_hi[u] * xsc[u], v_hi, acc_hi);
        }
    }
    if (!waited) pdl_wait();
    // columns: tig 0 and 1 hold the digits, tig 2 and 3 zeros
    acc_lo += __shfl_xor_sync(0xffffffffu, acc_lo, 1); acc_lo += __shfl_xor_sync(0xffffffffu, acc_lo, 2);
    acc_hi += __shfl_xor_sync(0xffffffffu, acc_hi, 1); acc_hi += __shfl_xor_sync(0xffffffffu, acc_hi, 2);
    if (tig == 0) { part[mtx][kw][gid] = acc_lo; part[mtx][kw][gid + 8] = acc_hi; }
    __syncthreads();
    if (threadIdx.x < 16 && R0 + threadIdx.x < M) {
        const uint32_t row = R0 + threadIdx.x;
        float s1 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < KS; i++) {
            s1 += part[0][i][threadIdx.x];
            if (SWIGLU) s3 += part[NM - 1][i][threadIdx.x];
        }
        float v = SWIGLU ? __fmul_rn(silu_ref(s1), s3) : s1;
        if (!SWIGLU && res) v = __fadd_rn(v, res[row]);
        y[row] = v;
    }
}

// scratch of the digits (experiment: one buffer per device — callers on several streams of one device would race)
static void q8_mma_scratch(uint32_t K, uint2 **bfrag, float **xs) {
    static thread_local int dev_of = -1;
    static thread_local uint2 *bf = nullptr;
    static thread_local float *sc = nullptr;
    static thread_local uint32_t cap = 0;
    int dev = 0;
    LB_CUDA(cudaGetDevice(&dev));
    if (dev != dev_of || (K >> 5) > cap) {
        const uint32_t nb = (K >> 5) > 1024 ? (K >> 5) : 1024;  // leaked on growth / device change: experiment only
        LB_CUDA(cudaMalloc(&bf, (size_t)nb * 32 * sizeof(uint2)));
        LB_CUDA(cudaMalloc(&sc, (size_t)nb * sizeof(float)));
        cap = nb; dev_of = dev;
    }
    *bfrag = bf; *xs = sc;
}

template <bool SWIGLU>
static void gemv_q8_mma(const int8_t *Q1, const float *D1, const int8_t *Q3, const float *D3, uint32_t M, uint32_t K, const float *x,
                        float *y, const float *res, cudaStream_t st) {
    uint2 *bfrag; float *xs;
    q8_mma_scratch(K, &bfrag, &xs);
    const uint32_t NB = K >> 5, tiles = M / 16;
    launch_pdl(q8_digits_kernel, dim3((NB + 3) / 4), dim3(128), 0, st, x, K, bfrag, xs);
    const uint32_t want = 148u * 16u;  // warps
    const uint32_t nm = SWIGLU ? 2 : 1;
    int ks = 2;
    while (ks < 16 && tiles * nm * ks < want) ks *= 2;
    if (SWIGLU && ks > 8) ks = 8;  // 2 * KS warps per block
#define LB_Q8M(KSV) launch_pdl(gemv_q8_mma_kernel<SWIGLU, KSV>, dim3(tiles), dim3(32 * KSV * nm), 0, st, Q1, D1, Q3, D3, M, K, \
                               (const uint2 *)bfrag, (const float *)xs, y, res)
    switch (ks) { case 2: LB_Q8M(2); break; case 4: LB_Q8M(4); break; case 8: LB_Q8M(8); break; default: if constexpr (!SWIGLU) { LB_Q8M(16); } else { LB_Q8M(8); } }
#undef LB_Q8M
}

// Resident blocks per SM the double-buffered kernel is compiled for: its grids must fit in ONE wave
// ([12288 x 4096] = 768 blocks -> 6 per SM; w1|w3 = 688 blocks -> 5; the K-split grids of 1024 blocks -> 7).
// (A bare minimum of 1 makes ptxas spend registers freely: 128 instead of 96 and a second wave for w1|w3.)
constexpr int q8_db_min_blocks(int nc, bool swiglu, int ksplit) {
    return nc == 1 ? (swiglu ? 5 : (ksplit == 1 ? 6 : 7)) : ((nc == 2 && !swiglu) ? 5 : 4);
}

// Double-buffered variant for 1-2 columns.  Every warp of these grids is resident at once and they start in
// phase, so the loop above pays (HBM latency + its share of the issue slots) serially once per batch: ~13 us
// of fixed cost on [12288 x 4096], which streams in 9 us.  Here the loads of batch i+1 are issued BEFORE the
// arithmetic of batch i.  To afford two batches in registers the 4 rows' block scales are not loaded by every
// lane (8 lanes share a block): lane j of each group of 8 loads the float4 of step j % U and the others
// fetch it with 4 shuffles.  k order and per-lane arithmetic are those of gemv_q8_kernel: same bits.
// (Rejected, profiles/README.md: a per-warp cp.async ring in shared memory — 1.5-2.5x slower.)
template <int NC, bool SWIGLU, int KSPLIT>
__global__ void __launch_bounds__(Q8_WARPS * 32, q8_db_min_blocks(NC, SWIGLU, KSPLIT))
gemv_q8_db_kernel(const int8_t *__restrict__ Q1, const float *__restrict__ D1, const int8_t *__restrict__ Q3,
                  const float *__restrict__ D3, uint32_t M, uint32_t K, const float *__restrict__ x, uint32_t ldx,
                  float *__restrict__ y, uint32_t ldy, const float *__restrict__ res) {
    constexpr int NM = SWIGLU ? 2 : 1;
    constexpr int U = Q8_UNROLL / NM;  // steps per batch (same register budget for W1 and W1 + W3)
    constexpr uint32_t STEP = 32 * U;
    __shared__ float part[KSPLIT > 1 ? Q8_WARPS : 1][NM][4][NC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g = KSPLIT > 1 ? blockIdx.x : blockIdx.x * Q8_WARPS + warp;  // row group
    const bool active = g < (M >> 2);
    const uint32_t K4 = K >> 2, KB = K >> 5;
    const uint32_t per = KSPLIT > 1 ? ((K4 / KSPLIT + 31) & ~31u) : K4;
    const uint32_t k4_begin = KSPLIT > 1 ? min(warp * per, K4) : 0, k4_end = KSPLIT > 1 ? min(k4_begin + per, K4) : K4;
    const uint4 *q1 = reinterpret_cast<const uint4 *>(Q1) + (size_t)(active ? g : 0) * K4;
    const float4 *d1 = reinterpret_cast<const float4 *>(D1) + (size_t)(active ? g : 0) * KB;
    const uint4 *q3 = SWIGLU ? reinterpret_cast<const uint4 *>(Q3) + (size_t)(active ? g : 0) * K4 : nullptr;
    const float4 *d3 = SWIGLU ? reinterpret_cast<const float4 *>(D3) + (size_t)(active ? g : 0) * KB : nullptr;
    const uint32_t su = (lane & 7) % U;              // the step whose scales this lane loads
    const uint32_t sq = (lane >> 3) * 8 + su * 32;   // k4 offset (within a batch) of that block's first lane
    float acc[NM][4][NC];
#pragma unroll
    for (int m = 0; m < NM; m++)
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < NC; c++) acc[m][r][c] = 0.f;

    uint4 w[2][U][NM];
    float4 s[2][NM];
    auto load_batch = [&](auto bc, uint32_t kb0) {  // kb0 = k4 of lane 0 at step 0 of the batch (multiple of 32)
        constexpr int b = decltype(bc)::value;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t k4 = kb0 + u * 32 + lane;
            const bool ok = active && k4 < k4_end;
            w[b][u][0] = ok ? ld_stream_u4(q1 + k4) : make_uint4(0, 0, 0, 0);
            if (SWIGLU) w[b][u][NM - 1] = ok ? ld_stream_u4(q3 + k4) : make_uint4(0, 0, 0, 0);
        }
        const uint32_t ks = kb0 + sq;  // first k4 of the block this lane fetches the scales of
        const bool oks = active && ks < k4_end;
        s[b][0] = oks ? __ldg(d1 + (ks >> 3)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (SWIGLU) s[b][NM - 1] = oks ? __ldg(d3 + (ks >> 3)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto compute_batch_stepwise = [&](auto bc, uint32_t kb0) {
        constexpr int b = decltype(bc)::value;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t k4 = kb0 + u * 32 + lane;
            const int src = (lane & ~7) | u;  // the lane of this group of 8 that holds step u's scales
            float sr[NM][4];
#pragma unroll
            for (int m = 0; m < NM; m++) {  // shuffles outside the k4 predicate: all 32 lanes take part
                sr[m][0] = __shfl_sync(0xffffffffu, s[b][m].x, src); sr[m][1] = __shfl_sync(0xffffffffu, s[b][m].y, src);
                sr[m][2] = __shfl_sync(0xffffffffu, s[b][m].z, src); sr[m][3] = __shfl_sync(0xffffffffu, s[b][m].w, src);
            }
            if (k4 < k4_end) {
                float4 xv[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) xv[c] = __ldg(reinterpret_cast<const float4 *>(x + (size_t)c * ldx) + k4);
#pragma unroll
                for (int m = 0; m < NM; m++) {
                    const uint32_t wr[4] = {w[b][u][m].x, w[b][u][m].y, w[b][u][m].z, w[b][u][m].w};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        float f[4];
                        unpack4(wr[r], f);
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            float t = f[0] * xv[c].x;
                            t = fmaf(f[1], xv[c].y, t); t = fmaf(f[2], xv[c].z, t); t = fmaf(f[3], xv[c].w, t);
                            acc[m][r][c] = fmaf(sr[m][r], t, acc[m][r][c]);
                        }
                    }
                }
            }
        }
    };
    auto compute_batch_hoisted = [&](auto bc, uint32_t kb0) {
        constexpr int b = decltype(bc)::value;
        // the batch's activation float4s first, back to back: loaded one step at a time, each step stalled on its
        // own L1 round trip (48 % of the stall samples in the ncu source view): 16.6 -> 14.2 us on [12288 x 4096]
        float4 xv[U][NC];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t k4 = kb0 + u * 32 + lane;
#pragma unroll
            for (int c = 0; c < NC; c++)
                xv[u][c] = k4 < k4_end ? __ldg(reinterpret_cast<const float4 *>(x + (size_t)c * ldx) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t k4 = kb0 + u * 32 + lane;
            const int src = (lane & ~7) | u;  // the lane of this group of 8 that holds step u's scales
            float sr[NM][4];
#pragma unroll
            for (int m = 0; m < NM; m++) {  // shuffles outside the k4 predicate: all 32 lanes take part
                sr[m][0] = __shfl_sync(0xffffffffu, s[b][m].x, src); sr[m][1] = __shfl_sync(0xffffffffu, s[b][m].y, src);
                sr[m][2] = __shfl_sync(0xffffffffu, s[b][m].z, src); sr[m][3] = __shfl_sync(0xffffffffu, s[b][m].w, src);
            }
            if (k4 < k4_end) {
#pragma unroll
                for (int m = 0; m < NM; m++) {
                    const uint32_t wr[4] = {w[b][u][m].x, w[b][u][m].y, w[b][u][m].z, w[b][u][m].w};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        float f[4];
                        unpack4(wr[r], f);
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            float t = f[0] * xv[u][c].x;
                            t = fmaf(f[1], xv[u][c].y, t); t = fmaf(f[2], xv[u][c].z, t); t = fmaf(f[3], xv[u][c].w, t);
                            acc[m][r][c] = fmaf(sr[m][r], t, acc[m][r][c]);
                        }
                    }
                }
            }
        }
    };
    // Whole-row single-column GEMV (qkv, lm_head) takes the hoisted form; the K-split and SwiGLU variants measured
    // slower with it (register pressure -> fewer resident blocks) and keep the per-step activation load.
    constexpr bool HOIST = NC == 1 && !SWIGLU && KSPLIT == 1;
    auto compute_batch = [&](auto bc, uint32_t kb0) {
        if constexpr (HOIST) compute_batch_hoisted(bc, kb0);
        else compute_batch_stepwise(bc, kb0);
    };
    using B0 = std::integral_constant<int, 0>;
    using B1 = std::integral_constant<int, 1>;
    pdl_launch_dependents();
    uint32_t kb0 = k4_begin;  // warp-uniform
    load_batch(B0{}, kb0);    // read-only weights first, then wait for the predecessor grid (PDL)
    if (kb0 + STEP < k4_end) load_batch(B1{}, kb0 + STEP);
    pdl_wait();
    while (kb0 < k4_end) {
        compute_batch(B0{}, kb0);
        if (kb0 + 2 * STEP < k4_end) load_batch(B0{}, kb0 + 2 * STEP);
        kb0 += STEP;
        if (kb0 >= k4_end) break;
        compute_batch(B1{}, kb0);
        if (kb0 + 2 * STEP < k4_end) load_batch(B1{}, kb0 + 2 * STEP);
        kb0 += STEP;
    }
#pragma unroll
    for (int m = 0; m < NM; m++)
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < NC; c++) acc[m][r][c] = warp_sum(acc[m][r][c]);
    if (KSPLIT > 1) {
        if (lane == 0) {
#pragma unroll
            for (int m = 0; m < NM; m++)
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < NC; c++) part[warp][m][r][c] = acc[m][r][c];
        }
        __syncthreads();
        if (warp != 0) return;
#pragma unroll
        for (int m = 0; m < NM; m++)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    float t = part[0][m][r][c];
                    for (int wv = 1; wv < Q8_WARPS; wv++) t += part[wv][m][r][c];
                    acc[m][r][c] = t;
                }
    }
    if (lane == 0 && active) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t row = g * 4 + r;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                float v = SWIGLU ? __fmul_rn(silu_ref(acc[0][r][c]), acc[NM - 1][r][c]) : acc[0][r][c];
                if (!SWIGLU && res) v = __fadd_rn(v, res[(size_t)c * ldy + row]);
                y[(size_t)c * ldy + row] = v;
            }
        }
    }
}

template <bool SWIGLU>
static void gemv_q8_dispatch(const int8_t *Q1, const float *D1, const int8_t *Q3, const float *D3, uint32_t M, uint32_t K,
                             const float *x, uint32_t ldx, uint32_t N, float *y, uint32_t ldy, const float *res, cudaStream_t st) {
    LB_CHECK(N >= 1 && N <= 8, "gemv_q8: N must be 1..8");
    LB_CHECK((K & 31) == 0 && (ldx & 3) == 0 && (M & 3) == 0, "gemv_q8: K must be a multiple of 32 and M of 4");
    const uint32_t groups = M / 4;
    const bool split = groups < 148u * 16u;  // few row groups (wo, w2): split K over the block's warps to fill the SMs
    static const bool use_mma = getenv("LB_Q8_MMA") != nullptr;  // experiment: int8 tensor cores (1 column, M % 16 == 0)
    if (use_mma && N == 1 && (M & 15) == 0) { gemv_q8_mma<SWIGLU>(Q1, D1, Q3, D3, M, K, x, y, res, st); return; }
    static const bool sync_loop = getenv("LB_Q8_SYNC") != nullptr;  // A/B aid: the single-buffered loop for every N
#define LB_Q8_LAUNCH(kern, n, ks)                                                                                                       \
    launch_pdl(kern<n, SWIGLU, ks>, dim3(ks > 1 ? groups : (groups + Q8_WARPS - 1) / Q8_WARPS), dim3(Q8_WARPS * 32), 0, st, Q1, D1, Q3, \
               D3, M, K, x, ldx, y, ldy, res)
#define LB_Q8_CASE_DB(n)                                                                                                                \
    case n:                                                                                                                             \
        if (sync_loop) { if (split) LB_Q8_LAUNCH(gemv_q8_kernel, n, 4); else LB_Q8_LAUNCH(gemv_q8_kernel, n, 1); }                     \
        else { if (split) LB_Q8_LAUNCH(gemv_q8_db_kernel, n, 4); else LB_Q8_LAUNCH(gemv_q8_db_kernel, n, 1); }                         \
        break;
#define LB_Q8_CASE(n)                                                                                                                   \
    case n:                                                                                                                             \
        if (split) LB_Q8_LAUNCH(gemv_q8_kernel, n, 4); else LB_Q8_LAUNCH(gemv_q8_kernel, n, 1);                                        \
        break;
    // 1-2 columns (single-sequence decode): double-buffered; 3-8 columns (pods) keep the registers for accumulators
    switch (N) { LB_Q8_CASE_DB(1) LB_Q8_CASE_DB(2) LB_Q8_CASE(3) LB_Q8_CASE(4) LB_Q8_CASE(5) LB_Q8_CASE(6) LB_Q8_CASE(7) default: LB_Q8_CASE(8) }
#undef LB_Q8_CASE
#undef LB_Q8_CASE_DB
#undef LB_Q8_LAUNCH
}
void gemv_q8(const int8_t *Q, const float *D, uint32_t M, uint32_t K, const float *x, uint32_t ldx, uint32_t N, float *y,
             uint32_t ldy, const float *residual, cudaStream_t st) {
    gemv_q8_dispatch<false>(Q, D, nullptr, nullptr, M, K, x, ldx, N, y, ldy, residual, st);
}
void gemv_q8_swiglu(const int8_t *Q1, const float *D1, const int8_t *Q3, const float *D3, uint32_t M, uint32_t K,
                    const float *x, uint32_t ldx, uint32_t N, float *act, uint32_t ldy, cudaStream_t st) {
    gemv_q8_dispatch<true>(Q1, D1, Q3, D3, M, K, x, ldx, N, act, ldy, nullptr, st);
}

// Prefill GEMM with the dequantisation fused into the shared-memory stage: the int8 tile and its
// scales are loaded, expanded to f32(d*q) while being written k-major into shared memory, then the
// same 128x64x16 FP32 register-tile loop as gemm_f32 runs.
constexpr int GM = 128, GN = 64, GK = 16;
__global__ void __launch_bounds__(256)
gemm_q8_kernel(const int8_t *__restrict__ Q, const float *__restrict__ D, uint32_t M, uint32_t K,
               const float *__restrict__ X, uint32_t ldx, uint32_t N, float *__restrict__ Y, uint32_t ldy,
               const float *__restrict__ res) {
    __shared__ float As[GK][GM + 4];
    __shared__ float Bs[GK][GN + 4];
    const uint32_t m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
    const int tid = threadIdx.x;
    const int tm = (tid & 15) * 8, tn = (tid >> 4) * 4;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (uint32_t k0 = 0; k0 < K; k0 += GK) {
        {   // 128 rows x 16 int8: thread t -> row t/2, 8 weights at (t%2)*8 (two 4-weight words of the interleaved layout)
            int r = tid >> 1, kq = (tid & 1) * 8;
            float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (m0 + r < M) {
                const uint32_t row = m0 + r;
                const uint32_t wa = *reinterpret_cast<const uint32_t *>(Q + q8_qoff(row, k0 + kq, K));
                const uint32_t wb = *reinterpret_cast<const uint32_t *>(Q + q8_qoff(row, k0 + kq + 4, K));
                const float dd = D[q8_doff(row, k0 + kq, K)];
                unpack4(wa, f); unpack4(wb, f + 4);
#pragma unroll
                for (int i = 0; i < 8; i++) f[i] = __fmul_rn(dd, f[i]);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) As[kq + i][r] = f[i];
        }
        {
            int r = tid >> 2, kq = (tid & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + r < N) v = *reinterpret_cast<const float4 *>(X + (size_t)(n0 + r) * ldx + k0 + kq);
            Bs[kq + 0][r] = v.x; Bs[kq + 1][r] = v.y; Bs[kq + 2][r] = v.z; Bs[kq + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; kk++) {
            float a[8], b[4];
            *reinterpret_cast<float4 *>(a) = *reinterpret_cast<const float4 *>(&As[kk][tm]);
            *reinterpret_cast<float4 *>(a + 4) = *reinterpret_cast<const float4 *>(&As[kk][tm + 4]);
            *reinterpret_cast<float4 *>(b) = *reinterpret_cast<const float4 *>(&Bs[kk][tn]);
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t n = n0 + tn + j;
        if (n >= N) continue;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t m = m0 + tm + i;
            if (m < M) {
                float v = acc[i][j];
                if (res) v = __fadd_rn(v, res[(size_t)n * ldy + m]);
                Y[(size_t)n * ldy + m] = v;
            }
        }
    }
}
void gemm_q8(const int8_t *Q, const float *D, uint32_t M, uint32_t K, const float *X, uint32_t ldx, uint32_t N, float *Y,
             uint32_t ldy, const float *residual, cudaStream_t st) {
    LB_CHECK((K & 31) == 0 && (ldx & 3) == 0 && (M & 3) == 0, "gemm_q8: K must be a multiple of 32 and M of 4");
    if (!N || !M) return;
    dim3 grid((M + GM - 1) / GM, (N + GN - 1) / GN);
    gemm_q8_kernel<<<grid, 256, 0, st>>>(Q, D, M, K, X, ldx, N, Y, ldy, residual);
    LB_LAUNCH_CHECK();
}

}  // namespace k
}  // namespace lb
