// kernels_q8.cu — block-quantised INT8 weights x FP32 activations (BASELINE config 3).
//
// The reference has no quantised path (only the enum space ml.go:85-94 and the block constants
// QK = 32, ml.go:24,123-124; INT8 is an unchecked roadmap item, README.md:45), so the format is
// defined here consistently with those constants:
//   block = 32 consecutive weights of one row (along K);  d = max|w| / 127  (FP32);
//   q_i = rint(w_i / d) clamped to [-127, 127] (FP32 divide, round-half-even); d == 0 -> q = 0.
//   36 bytes per 32 weights, stored as two planes: q[M][K] int8 and d[M][K/32] float
//   (same bytes as the interleaved block, but 16-byte aligned for vector loads).
// Parity target: the FP32 path on the dequantised weights f32(d * q_i).  The kernels compute
// f32(d*q_i) explicitly (one rounding) and then FMA, i.e. the same products as the target.
#include "common.cuh"
#include "kernels.cuh"

namespace lb {
namespace k {

__global__ void quantize_q8_kernel(const float *__restrict__ W, int8_t *__restrict__ q, float *__restrict__ d, size_t nblocks) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const float4 *src = reinterpret_cast<const float4 *>(W + b * 32);
    float v[32];
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float4 t = src[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
    }
    const float dd = __fdiv_rn(amax, 127.0f);
    d[b] = dd;
    int8_t out[32];
#pragma unroll
    for (int i = 0; i < 32; i++) {
        int r = dd > 0.f ? __float2int_rn(__fdiv_rn(v[i], dd)) : 0;
        r = max(-127, min(127, r));
        out[i] = (int8_t)r;
    }
    int4 *dst = reinterpret_cast<int4 *>(q + b * 32);
    dst[0] = *reinterpret_cast<int4 *>(out);
    dst[1] = *reinterpret_cast<int4 *>(out + 16);
}
void quantize_q8(const float *W, int8_t *q, float *d, size_t nelem, cudaStream_t st) {
    LB_CHECK(nelem % 32 == 0, "quantize_q8: element count must be a multiple of 32");
    size_t nb = nelem / 32;
    if (!nb) return;
    quantize_q8_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(W, q, d, nb);
    LB_LAUNCH_CHECK();
}

__global__ void dequantize_q8_kernel(const int8_t *__restrict__ q, const float *__restrict__ d, float *__restrict__ out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __fmul_rn(d[i >> 5], (float)q[i]);
}
void dequantize_q8(const int8_t *q, const float *d, float *out, size_t nelem, cudaStream_t st) {
    if (!nelem) return;
    size_t blocks = (nelem + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    dequantize_q8_kernel<<<(unsigned)blocks, 256, 0, st>>>(q, d, out, nelem);
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

// Decode GEMV.  One warp per output row (or per w1/w3 row pair); a lane owns 4 consecutive weights
// per step (one coalesced 128-byte warp request of int8 against one coalesced 512-byte request of
// FP32 activations — the activations are the wider stream here), Q8_UNROLL requests in flight.
constexpr int Q8_WARPS = 4;
constexpr int Q8_UNROLL = 8;

template <int NC, bool SWIGLU>
__global__ void __launch_bounds__(Q8_WARPS * 32)
gemv_q8_kernel(const int8_t *__restrict__ Q1, const float *__restrict__ D1, const int8_t *__restrict__ Q3,
               const float *__restrict__ D3, uint32_t M, uint32_t K, const float *__restrict__ x, uint32_t ldx,
               float *__restrict__ y, uint32_t ldy, const float *__restrict__ res) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row = blockIdx.x * Q8_WARPS + warp;
    if (row >= M) return;
    const uint32_t *q1 = reinterpret_cast<const uint32_t *>(Q1 + (size_t)row * K);
    const float *d1 = D1 + (size_t)row * (K >> 5);
    const uint32_t *q3 = SWIGLU ? reinterpret_cast<const uint32_t *>(Q3 + (size_t)row * K) : nullptr;
    const float *d3 = SWIGLU ? D3 + (size_t)row * (K >> 5) : nullptr;
    float a1[NC], a3[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) a1[c] = a3[c] = 0.f;
    const uint32_t K4 = K >> 2;
    uint32_t w1[Q8_UNROLL], w3[Q8_UNROLL];
    float s1[Q8_UNROLL], s3[Q8_UNROLL];
    auto load_batch = [&](uint32_t kk) {
#pragma unroll
        for (int u = 0; u < Q8_UNROLL; u++) {
            uint32_t k4 = kk + u * 32;
            bool ok = k4 < K4;
            w1[u] = ok ? __ldg(q1 + k4) : 0x0u;
            s1[u] = ok ? __ldg(d1 + (k4 >> 3)) : 0.f;
            if (SWIGLU) {
                w3[u] = ok ? __ldg(q3 + k4) : 0x0u;
                s3[u] = ok ? __ldg(d3 + (k4 >> 3)) : 0.f;
            }
        }
    };
    pdl_launch_dependents();
    load_batch(lane);  // weights and scales are read-only: issue them before waiting on the predecessor grid (PDL)
    pdl_wait();
    for (uint32_t kk = lane; kk < K4;) {
#pragma unroll
        for (int u = 0; u < Q8_UNROLL; u++) {
            uint32_t k4 = kk + u * 32;
            if (k4 < K4) {
                float f1[4], f3[4];
                unpack4(w1[u], f1);
#pragma unroll
                for (int i = 0; i < 4; i++) f1[i] = __fmul_rn(s1[u], f1[i]);
                if (SWIGLU) {
                    unpack4(w3[u], f3);
#pragma unroll
                    for (int i = 0; i < 4; i++) f3[i] = __fmul_rn(s3[u], f3[i]);
                }
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    float4 xv = __ldg(reinterpret_cast<const float4 *>(x + (size_t)c * ldx) + k4);
                    a1[c] = fmaf(f1[0], xv.x, a1[c]); a1[c] = fmaf(f1[1], xv.y, a1[c]);
                    a1[c] = fmaf(f1[2], xv.z, a1[c]); a1[c] = fmaf(f1[3], xv.w, a1[c]);
                    if (SWIGLU) {
                        a3[c] = fmaf(f3[0], xv.x, a3[c]); a3[c] = fmaf(f3[1], xv.y, a3[c]);
                        a3[c] = fmaf(f3[2], xv.z, a3[c]); a3[c] = fmaf(f3[3], xv.w, a3[c]);
                    }
                }
            }
        }
        kk += 32 * Q8_UNROLL;
        if (kk < K4) load_batch(kk);
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {
        a1[c] = warp_sum(a1[c]);
        if (SWIGLU) a3[c] = warp_sum(a3[c]);
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < NC; c++) {
            float v = SWIGLU ? __fmul_rn(silu_ref(a1[c]), a3[c]) : a1[c];
            if (!SWIGLU && res) v = __fadd_rn(v, res[(size_t)c * ldy + row]);
            y[(size_t)c * ldy + row] = v;
        }
    }
}

template <bool SWIGLU>
static void gemv_q8_dispatch(const int8_t *Q1, const float *D1, const int8_t *Q3, const float *D3, uint32_t M, uint32_t K,
                             const float *x, uint32_t ldx, uint32_t N, float *y, uint32_t ldy, const float *res, cudaStream_t st) {
    LB_CHECK(N >= 1 && N <= 8, "gemv_q8: N must be 1..8");
    LB_CHECK((K & 31) == 0 && (ldx & 3) == 0, "gemv_q8: K must be a multiple of 32");
    unsigned grid = (M + Q8_WARPS - 1) / Q8_WARPS;
#define LB_Q8_CASE(n) case n: launch_pdl(gemv_q8_kernel<n, SWIGLU>, dim3(grid), dim3(Q8_WARPS * 32), 0, st, Q1, D1, Q3, D3, M, K, x, ldx, y, ldy, res); break;
    switch (N) { LB_Q8_CASE(1) LB_Q8_CASE(2) LB_Q8_CASE(3) LB_Q8_CASE(4) LB_Q8_CASE(5) LB_Q8_CASE(6) LB_Q8_CASE(7) default: LB_Q8_CASE(8) }
#undef LB_Q8_CASE
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
    const uint32_t KB = K >> 5;
    for (uint32_t k0 = 0; k0 < K; k0 += GK) {
        {   // 128 rows x 16 int8: thread t -> row t/2, 8 weights at (t%2)*8
            int r = tid >> 1, kq = (tid & 1) * 8;
            float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (m0 + r < M) {
                const uint2 w = *reinterpret_cast<const uint2 *>(Q + (size_t)(m0 + r) * K + k0 + kq);
                const float dd = D[(size_t)(m0 + r) * KB + ((k0 + kq) >> 5)];
                unpack4(w.x, f); unpack4(w.y, f + 4);
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
    LB_CHECK((K & 31) == 0 && (ldx & 3) == 0, "gemm_q8: K must be a multiple of 32");
    if (!N || !M) return;
    dim3 grid((M + GM - 1) / GM, (N + GN - 1) / GN);
    gemm_q8_kernel<<<grid, 256, 0, st>>>(Q, D, M, K, X, ldx, N, Y, ldy, residual);
    LB_LAUNCH_CHECK();
}

}  // namespace k
}  // namespace lb
