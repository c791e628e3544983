// kernels_attn.cu — causal attention over the FP32 KV cache: the reference's
//   KQ = MulMat(K, Q); Scale(1/sqrt(hd)); DiagMaskInf(past); SoftMax; KQV = MulMat(V^T, P)
// chain (pkg/llama/llama.go:300-333) fused into one kernel.  No V^T copy is materialised
// (the reference re-transposes the whole layer's V cache every call, llama.go:315-322).
// Numerics: FP32 dot, FP32 multiply by f32(1/sqrt(hd)), FP32 max, e = f32(exp(f64(s - max))),
// p = e * f32(1/sum) (ml.go:2472-2499), FP32 P·V.  The sums are tree-ordered.
#include "common.cuh"
#include "kernels.cuh"

namespace lb {
namespace k {

constexpr int ATT_THREADS = 256;

// grid (heads, N): one CTA per (head, query).  Dynamic smem: scores[T].
__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const float *__restrict__ q, uint32_t ldq, const float *__restrict__ Kc, const float *__restrict__ Vc,
                 float *__restrict__ out, const uint32_t *__restrict__ past_dev, uint32_t max_T, uint32_t dim,
                 uint32_t hd, float scale) {
    extern __shared__ float sm[];
    const uint32_t past = *past_dev;
    __shared__ float red[ATT_THREADS / 32];
    __shared__ float s_bcast;
    const uint32_t h = blockIdx.x, n = blockIdx.y;
    const uint32_t Tn = past + n + 1;  // causal: keys 0 .. past+n   (DiagMaskInf, ml.go:2399-2408)
    float *scores = sm;                // [Tn]
    float *part = sm + ((max_T + 3) & ~3u);  // [groups][hd] partial P·V sums
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *qh = q + (size_t)n * ldq + (size_t)h * hd;

    // ---- scores
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool act = (uint32_t)(lane * 4) < hd;
    if (act) qv = *reinterpret_cast<const float4 *>(qh + lane * 4);
    for (uint32_t t = warp; t < Tn; t += ATT_THREADS / 32) {
        float d = 0.f;
        if (act) {
            float4 kv = *reinterpret_cast<const float4 *>(Kc + (size_t)t * dim + (size_t)h * hd + lane * 4);
            d = fmaf(kv.x, qv.x, d); d = fmaf(kv.y, qv.y, d); d = fmaf(kv.z, qv.z, d); d = fmaf(kv.w, qv.w, d);
        }
        if (hd > 4) d = warp_sum(d);
        if (lane == 0) scores[t] = __fmul_rn(d, scale);
    }
    __syncthreads();
    // ---- softmax
    float m = -INFINITY;
    for (uint32_t t = threadIdx.x; t < Tn; t += ATT_THREADS) m = fmaxf(m, scores[t]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = red[0];
        for (int i = 1; i < ATT_THREADS / 32; i++) t = fmaxf(t, red[i]);
        s_bcast = t;
    }
    __syncthreads();
    m = s_bcast;
    float sum = 0.f;
    for (uint32_t t = threadIdx.x; t < Tn; t += ATT_THREADS) {
        float e = (float)exp((double)__fsub_rn(scores[t], m));
        scores[t] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < ATT_THREADS / 32; i++) t += red[i];
        s_bcast = __fdiv_rn(1.0f, t);
    }
    __syncthreads();
    const float inv = s_bcast;
    // ---- P·V: thread (g, d) accumulates keys t = g, g+G, ...
    const uint32_t G = ATT_THREADS / hd;  // hd in {32, 64, 128}: 8 / 4 / 2 groups
    const uint32_t g = threadIdx.x / hd, d = threadIdx.x % hd;
    float acc = 0.f;
    if (g < G) {
        const float *vp = Vc + (size_t)h * hd + d;
        for (uint32_t t = g; t < Tn; t += G) acc = fmaf(vp[(size_t)t * dim], __fmul_rn(scores[t], inv), acc);
        part[g * hd + d] = acc;
    }
    __syncthreads();
    if (threadIdx.x < hd) {
        float r = 0.f;
        for (uint32_t i = 0; i < G; i++) r += part[i * hd + threadIdx.x];
        out[(size_t)n * dim + (size_t)h * hd + threadIdx.x] = r;
    }
}

void attention(const float *q, uint32_t ldq, const float *Kc, const float *Vc, float *out, uint32_t N,
               const uint32_t *past_dev, uint32_t max_T, uint32_t dim, uint32_t heads, cudaStream_t st) {
    const uint32_t hd = dim / heads;
    LB_CHECK(hd == 32 || hd == 64 || hd == 128, "attention: head dim must be 32, 64 or 128");
    const uint32_t T = max_T;
    size_t smem = (((size_t)T + 3) & ~(size_t)3) * sizeof(float) + (size_t)ATT_THREADS * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        LB_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    LB_CHECK(smem <= 200 * 1024, "attention: context too long for the single-pass kernel");
    float scale = (float)(1.0 / sqrt((double)dim / (double)heads));  // llama.go:306
    attention_kernel<<<dim3(heads, N), ATT_THREADS, smem, st>>>(q, ldq, Kc, Vc, out, past_dev, max_T, dim, hd, scale);
    LB_LAUNCH_CHECK();
}

}  // namespace k
}  // namespace lb
