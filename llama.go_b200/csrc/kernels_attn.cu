// kernels_attn.cu — causal attention over the FP32 KV cache: the reference's
//   KQ = MulMat(K, Q); Scale(1/sqrt(hd)); DiagMaskInf(past); SoftMax; KQV = MulMat(V^T, P)
// chain (pkg/llama/llama.go:300-333) fused into one kernel.  No V^T copy is materialised
// (the reference re-transposes the whole layer's V cache every call, llama.go:315-322).
// Numerics: FP32 dot, FP32 multiply by f32(1/sqrt(hd)), FP32 max, e = f32(exp(f64(s - max))),
// p = e * f32(1/sum) (ml.go:2472-2499), FP32 P·V.  The sums are tree-ordered.
#include <stdlib.h>

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

// ------------------------------------------------------------------------------------------
// Decode attention (N == 1): split the T cached positions over SPLITS CTAs per head so that the
// whole chip streams the layer's K/V slab (2*T*dim*4 bytes) instead of one CTA per head.
// Each CTA: scores for its key range, local max m, e = f32(exp(f64(s - m))), local sum l, partial
// O = sum_t e_t * V[t]; the last CTA to finish a head (atomic ticket) merges the partials:
// M = max m_s, L = sum l_s * w_s, out = (sum O_s * w_s) * f32(1/L), w_s = f32(exp(f64(m_s - M))).
// Versus the reference's single-pass softmax this reassociates the sums and applies the
// normaliser after P·V — a few 1e-7 relative.
// ------------------------------------------------------------------------------------------
constexpr int DEC_THREADS = 128;
constexpr int DEC_MAX_SPLITS = 32;

template <int HD>
__global__ void __launch_bounds__(DEC_THREADS)
attention_decode_kernel(const float *__restrict__ q, const float *__restrict__ Kc, const float *__restrict__ Vc,
                        float *__restrict__ out, const uint32_t *__restrict__ past_dev, uint32_t dim,
                        float scale, float *__restrict__ part_o, float *__restrict__ part_ml,
                        unsigned int *__restrict__ tickets, uint32_t chunk_cap, PodPtrs pods) {
    extern __shared__ float sm[];  // scores[chunk_cap]
    __shared__ float red[DEC_THREADS / 32];
    __shared__ float s_bcast;
    __shared__ unsigned int s_ticket;
    const uint32_t h = blockIdx.x, sp = blockIdx.y, S = gridDim.y;
    pdl_launch_dependents();
    pdl_wait();
    if (pods.K) {  // pod batch: sequence b = blockIdx.z has its own cache, position and rows
        const uint32_t b = blockIdx.z, H = gridDim.x;
        Kc = pods.K[b] + pods.layer_off; Vc = pods.V[b] + pods.layer_off;
        past_dev = pods.pasts + b;
        q += (size_t)b * pods.ldq; out += (size_t)b * pods.ldo;
        part_o += (size_t)b * H * DEC_MAX_SPLITS * HD; part_ml += (size_t)b * H * DEC_MAX_SPLITS * 2; tickets += (size_t)b * H;
    }
    const uint32_t Tn = *past_dev + 1;
    const uint32_t chunk = min((Tn + S - 1) / S, chunk_cap);
    const uint32_t t0 = min(sp * chunk, Tn), t1 = min(t0 + chunk, Tn);
    const uint32_t nk = t1 - t0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int LANES = HD / 4;  // lanes that hold a float4 of the head
    const float *qh = q + (size_t)h * HD;
    const float *kbase = Kc + (size_t)h * HD, *vbase = Vc + (size_t)h * HD;

    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < LANES) qv = *reinterpret_cast<const float4 *>(qh + lane * 4);
    // ---- scores: warp w takes keys t0 + w, t0 + w + NW, ...; AU keys in flight per warp
    constexpr int NW = DEC_THREADS / 32;
    constexpr int AU = 8;
    for (uint32_t i = warp; i < nk; i += NW * AU) {
        float4 kv[AU];
#pragma unroll
        for (int u = 0; u < AU; u++) {
            uint32_t ii = i + u * NW;
            kv[u] = (ii < nk && lane < LANES) ? ld_stream_f4(kbase + (size_t)(t0 + ii) * dim + lane * 4)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < AU; u++) {
            uint32_t ii = i + u * NW;
            float d = kv[u].x * qv.x;
            d = fmaf(kv[u].y, qv.y, d); d = fmaf(kv[u].z, qv.z, d); d = fmaf(kv[u].w, qv.w, d);
            d = warp_sum(d);
            if (lane == 0 && ii < nk) sm[ii] = __fmul_rn(d, scale);
        }
    }
    // ---- the first AU V rows of this thread do not depend on the scores: fetch them now, under the softmax.
    // Thread (kg, dl) takes keys kg, kg + KG, ... for the 4 dims of float4 lane dl.
    constexpr int KG = DEC_THREADS / LANES;
    const uint32_t kg = threadIdx.x / LANES, dl = threadIdx.x % LANES;
    float4 vf[AU];
#pragma unroll
    for (int u = 0; u < AU; u++) {
        const uint32_t key = kg + u * KG;
        vf[u] = key < nk ? ld_stream_f4(vbase + (size_t)(t0 + key) * dim + dl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    // ---- local softmax statistics
    float m = -INFINITY;
    for (uint32_t i = threadIdx.x; i < nk; i += DEC_THREADS) m = fmaxf(m, sm[i]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = red[0];
        for (int i = 1; i < NW; i++) t = fmaxf(t, red[i]);
        s_bcast = t;
    }
    __syncthreads();
    m = s_bcast;
    float l = 0.f;
    for (uint32_t i = threadIdx.x; i < nk; i += DEC_THREADS) {
        float e = (float)exp((double)__fsub_rn(sm[i], m));
        sm[i] = e;
        l += e;
    }
    l = warp_sum(l);
    __syncthreads();
    if (lane == 0) red[warp] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < NW; i++) t += red[i];
        part_ml[((size_t)h * S + sp) * 2 + 0] = m;
        part_ml[((size_t)h * S + sp) * 2 + 1] = t;
    }
    // ---- partial P·V: sequential over this thread's keys, then over the key groups (fixed order)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t base = 0; base < nk; base += KG * AU) {
        if (base) {
#pragma unroll
            for (int u = 0; u < AU; u++) {
                const uint32_t key = base + kg + u * KG;
                vf[u] = key < nk ? ld_stream_f4(vbase + (size_t)(t0 + key) * dim + dl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const uint32_t key = base + kg + u * KG;
            if (key < nk) {
                const float sc = sm[key];
                acc.x = fmaf(vf[u].x, sc, acc.x); acc.y = fmaf(vf[u].y, sc, acc.y);
                acc.z = fmaf(vf[u].z, sc, acc.z); acc.w = fmaf(vf[u].w, sc, acc.w);
            }
        }
    }
    {
        __shared__ float4 pv[DEC_THREADS];  // [kg][dl]
        pv[threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x < HD) {
            const float *pvf = reinterpret_cast<const float *>(pv);
            float r = 0.f;
            for (int i = 0; i < KG; i++) r += pvf[i * HD + threadIdx.x];
            part_o[((size_t)h * S + sp) * HD + threadIdx.x] = r;
        }
    }
    // ---- ticket: the last CTA of this head merges
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[h], 1u);
    __syncthreads();
    if (s_ticket != S - 1) return;
    __threadfence();
    // Merge of the S <= 32 partials.  The statistics are fetched by S lanes at once and every thread's S partial
    // outputs 8 at a time: a loop of dependent L2 reads (3 per split) cost ~0.4 us per split, more than the
    // attention itself.  The accumulation order over the splits stays sequential, so the bits do not change.
    __shared__ float s_w[DEC_MAX_SPLITS], s_l[DEC_MAX_SPLITS];
    if (warp == 0) {
        float ms = -INFINITY, ls = 0.f;
        if ((uint32_t)lane < S) {
            ms = __ldcg(&part_ml[((size_t)h * S + lane) * 2]);
            ls = __ldcg(&part_ml[((size_t)h * S + lane) * 2 + 1]);
        }
        const float M = warp_max(ms);
        // an empty split has m = -inf, l = 0, o = 0: its weight is exp(-inf) = 0 and it adds exact zeros
        s_w[lane] = ls > 0.f ? (float)exp((double)__fsub_rn(ms, M)) : 0.f;
        s_l[lane] = ls;
    }
    __syncthreads();
    if (threadIdx.x < HD) {
        float L = 0.f, o = 0.f;
        const float *po = part_o + (size_t)h * S * HD + threadIdx.x;
        for (uint32_t s0 = 0; s0 < S; s0 += 8) {
            float pv8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) pv8[u] = s0 + u < S ? __ldcg(po + (size_t)(s0 + u) * HD) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (s0 + u < S && s_l[s0 + u] > 0.f) {
                    L = fmaf(s_l[s0 + u], s_w[s0 + u], L);
                    o = fmaf(pv8[u], s_w[s0 + u], o);
                }
            }
        }
        out[(size_t)h * HD + threadIdx.x] = __fmul_rn(o, __fdiv_rn(1.0f, L));
    }
    if (threadIdx.x == 0) tickets[h] = 0;  // ready for the next launch / graph replay
}

uint32_t attention_decode_splits(uint32_t max_T) {
    uint32_t s = (max_T + 31) / 32;
    if (s < 1) s = 1;
    if (s > DEC_MAX_SPLITS) s = DEC_MAX_SPLITS;
    return s;
}
size_t attention_decode_scratch_floats(uint32_t heads, uint32_t hd) {
    // part_o [H][S][hd] + part_ml [H][S][2] + tickets [H] (as uint32)
    return (size_t)heads * DEC_MAX_SPLITS * (hd + 2) + heads;
}

static void attention_decode_launch(const float *q, const float *Kc, const float *Vc, float *out, const uint32_t *past_dev,
                                    uint32_t max_T, uint32_t dim, uint32_t heads, float *scratch, uint32_t B, const PodPtrs &pods,
                                    cudaStream_t st);

void attention_decode(const float *q, const float *Kc, const float *Vc, float *out, const uint32_t *past_dev,
                      uint32_t max_T, uint32_t dim, uint32_t heads, float *scratch, cudaStream_t st) {
    attention_decode_launch(q, Kc, Vc, out, past_dev, max_T, dim, heads, scratch, 1, PodPtrs{}, st);
}
void attention_decode_pods(const float *q, float *out, uint32_t B, const PodPtrs &pods, uint32_t max_T, uint32_t dim,
                           uint32_t heads, float *scratch, cudaStream_t st) {
    attention_decode_launch(q, nullptr, nullptr, out, nullptr, max_T, dim, heads, scratch, B, pods, st);
}

static void attention_decode_launch(const float *q, const float *Kc, const float *Vc, float *out, const uint32_t *past_dev,
                                    uint32_t max_T, uint32_t dim, uint32_t heads, float *scratch, uint32_t B, const PodPtrs &pods,
                                    cudaStream_t st) {
    const uint32_t hd = dim / heads;
    LB_CHECK(hd == 32 || hd == 64 || hd == 128, "attention: head dim must be 32, 64 or 128");
    const uint32_t S = attention_decode_splits(max_T);
    const uint32_t chunk_cap = (max_T + S - 1) / S;
    // scratch layout (B = number of pods, 1 for a single sequence): part_o [B][H][32][hd] | part_ml [B][H][32][2] | tickets [B][H]
    float *part_o = scratch;
    float *part_ml = part_o + (size_t)B * heads * DEC_MAX_SPLITS * hd;
    unsigned int *tickets = reinterpret_cast<unsigned int *>(part_ml + (size_t)B * heads * DEC_MAX_SPLITS * 2);
    float scale = (float)(1.0 / sqrt((double)dim / (double)heads));  // llama.go:306
    size_t smem = (size_t)chunk_cap * sizeof(float);
    LB_CHECK(smem <= 40 * 1024, "attention_decode: context too long");
    dim3 grid(heads, S, B);
    if (hd == 128)
        launch_pdl(attention_decode_kernel<128>, grid, dim3(DEC_THREADS), smem, st, q, Kc, Vc, out, past_dev, dim, scale, part_o, part_ml, tickets, chunk_cap, pods);
    else if (hd == 64)
        launch_pdl(attention_decode_kernel<64>, grid, dim3(DEC_THREADS), smem, st, q, Kc, Vc, out, past_dev, dim, scale, part_o, part_ml, tickets, chunk_cap, pods);
    else
        launch_pdl(attention_decode_kernel<32>, grid, dim3(DEC_THREADS), smem, st, q, Kc, Vc, out, past_dev, dim, scale, part_o, part_ml, tickets, chunk_cap, pods);
}

// ------------------------------------------------------------------------------------------
// Prefill attention, tiled (N > 1): one CTA per (head, 16 consecutive queries).  The single-pass kernel above
// re-reads every query's whole K/V prefix from L2 (O(N*T) traffic: 2.4 GB per 7B layer at N = 384); here a
// 64-key tile of K and V is staged ONCE in shared memory for the 16 queries of the CTA (two per warp),
// scores by FP32 dots out of shared memory, online softmax per query (running max m, sum l and output acc,
// rescaled by f32(exp(f64(m - m'))) when the maximum moves; the terms e = f32(exp(f64(s - m'))) as in
// ml.go:2472-2499), P.V from the staged V tile, out = acc * f32(1/l).  Versus the reference's two-pass row softmax
// this reassociates the sums (a few 1e-7 relative).
// ------------------------------------------------------------------------------------------
constexpr int ATP_THREADS = 256;   // 8 warps x 2 queries
constexpr int ATP_QT = 16;         // queries per CTA
constexpr int ATP_KT = 64;         // keys per tile: lane l scores keys l and l + 32

template <int HD>
__global__ void __launch_bounds__(ATP_THREADS)
attention_tiled_kernel(const float *__restrict__ q, uint32_t ldq, const float *__restrict__ Kc, const float *__restrict__ Vc,
                       float *__restrict__ out, const uint32_t *__restrict__ past_dev, uint32_t N, uint32_t dim, float scale) {
    constexpr int KP = HD + 4;                 // K tile row pitch: 4 words mod 32 -> the 8 lanes of an LDS.128 phase hit 8 bank groups
    constexpr int D4 = HD / 4;
    extern __shared__ float smt[];
    float *Ks = smt;                           // [ATP_KT][KP]
    float *Vs = Ks + ATP_KT * KP;              // [ATP_KT][HD]
    float *Qs = Vs + ATP_KT * HD;              // [ATP_QT][HD]
    float *Ps = Qs + ATP_QT * HD;              // [ATP_QT][ATP_KT]
    const uint32_t past = *past_dev;
    const uint32_t h = blockIdx.x, q0 = blockIdx.y * ATP_QT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nq = min((uint32_t)ATP_QT, N - q0);
    const uint32_t Tmax = past + q0 + nq;      // keys 0 .. Tmax-1 are visible to the last query of the tile (DiagMaskInf, ml.go:2399-2408)
    for (uint32_t i = threadIdx.x; i < ATP_QT * D4; i += ATP_THREADS) {
        const uint32_t qi = i / D4, d4 = i % D4;
        reinterpret_cast<float4 *>(Qs)[i] = qi < nq ? *reinterpret_cast<const float4 *>(q + (size_t)(q0 + qi) * ldq + (size_t)h * HD + d4 * 4)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const uint32_t qa = warp * 2, qb = warp * 2 + 1;           // this warp's two queries
    const uint32_t ta = past + q0 + qa, tb = past + q0 + qb;   // last visible key of each
    float ma = -INFINITY, mb = -INFINITY, la = 0.f, lb = 0.f;
    float4 oa = make_float4(0.f, 0.f, 0.f, 0.f), ob = oa;      // lane l owns output dims 4l .. 4l+3 (lanes >= D4 idle for small heads)
    for (uint32_t k0 = 0; k0 < Tmax; k0 += ATP_KT) {
        __syncthreads();                                       // previous tile fully consumed (and Qs written, first trip)
        const uint32_t nk = min((uint32_t)ATP_KT, Tmax - k0);
        for (uint32_t i = threadIdx.x; i < ATP_KT * D4; i += ATP_THREADS) {
            const uint32_t kj = i / D4, d4 = i % D4;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (kj < nk) {
                kv = *reinterpret_cast<const float4 *>(Kc + (size_t)(k0 + kj) * dim + (size_t)h * HD + d4 * 4);
                vv = *reinterpret_cast<const float4 *>(Vc + (size_t)(k0 + kj) * dim + (size_t)h * HD + d4 * 4);
            }
            *reinterpret_cast<float4 *>(Ks + kj * KP + d4 * 4) = kv;
            *reinterpret_cast<float4 *>(Vs + kj * HD + d4 * 4) = vv;
        }
        __syncthreads();
        if (qa < nq) {
            // ---- scores of keys k0 + lane and k0 + lane + 32 for both queries
            float sa0 = 0.f, sa1 = 0.f, sb0 = 0.f, sb1 = 0.f;
            const float4 *k0p = reinterpret_cast<const float4 *>(Ks + lane * KP), *k1p = reinterpret_cast<const float4 *>(Ks + (lane + 32) * KP);
            const float4 *qap = reinterpret_cast<const float4 *>(Qs + qa * HD), *qbp = reinterpret_cast<const float4 *>(Qs + qb * HD);
#pragma unroll 8
            for (int d4 = 0; d4 < D4; d4++) {
                const float4 ka = k0p[d4], kb = k1p[d4], xa = qap[d4], xb = qbp[d4];
                sa0 = fmaf(ka.x, xa.x, sa0); sa0 = fmaf(ka.y, xa.y, sa0); sa0 = fmaf(ka.z, xa.z, sa0); sa0 = fmaf(ka.w, xa.w, sa0);
                sa1 = fmaf(kb.x, xa.x, sa1); sa1 = fmaf(kb.y, xa.y, sa1); sa1 = fmaf(kb.z, xa.z, sa1); sa1 = fmaf(kb.w, xa.w, sa1);
                sb0 = fmaf(ka.x, xb.x, sb0); sb0 = fmaf(ka.y, xb.y, sb0); sb0 = fmaf(ka.z, xb.z, sb0); sb0 = fmaf(ka.w, xb.w, sb0);
                sb1 = fmaf(kb.x, xb.x, sb1); sb1 = fmaf(kb.y, xb.y, sb1); sb1 = fmaf(kb.z, xb.z, sb1); sb1 = fmaf(kb.w, xb.w, sb1);
            }
            const uint32_t key0 = k0 + lane, key1 = k0 + lane + 32;
            // Scale (llama.go:303-307) then the causal mask
            sa0 = key0 <= ta ? __fmul_rn(sa0, scale) : -INFINITY; sa1 = key1 <= ta ? __fmul_rn(sa1, scale) : -INFINITY;
            sb0 = (qb < nq && key0 <= tb) ? __fmul_rn(sb0, scale) : -INFINITY; sb1 = (qb < nq && key1 <= tb) ? __fmul_rn(sb1, scale) : -INFINITY;
            // ---- online softmax, query a
            {
                const float mt = warp_max(fmaxf(sa0, sa1));
                const float mn = fmaxf(ma, mt);
                if (mn != -INFINITY) {
                    const float alpha = ma == -INFINITY ? 0.f : (float)exp((double)__fsub_rn(ma, mn));
                    const float e0 = sa0 == -INFINITY ? 0.f : (float)exp((double)__fsub_rn(sa0, mn));
                    const float e1 = sa1 == -INFINITY ? 0.f : (float)exp((double)__fsub_rn(sa1, mn));
                    Ps[qa * ATP_KT + lane] = e0; Ps[qa * ATP_KT + lane + 32] = e1;
                    la = fmaf(la, alpha, warp_sum(e0 + e1));
                    oa.x *= alpha; oa.y *= alpha; oa.z *= alpha; oa.w *= alpha;
                    ma = mn;
                } else { Ps[qa * ATP_KT + lane] = 0.f; Ps[qa * ATP_KT + lane + 32] = 0.f; }
            }
            if (qb < nq) {
                const float mt = warp_max(fmaxf(sb0, sb1));
                const float mn = fmaxf(mb, mt);
                if (mn != -INFINITY) {
                    const float alpha = mb == -INFINITY ? 0.f : (float)exp((double)__fsub_rn(mb, mn));
                    const float e0 = sb0 == -INFINITY ? 0.f : (float)exp((double)__fsub_rn(sb0, mn));
                    const float e1 = sb1 == -INFINITY ? 0.f : (float)exp((double)__fsub_rn(sb1, mn));
                    Ps[qb * ATP_KT + lane] = e0; Ps[qb * ATP_KT + lane + 32] = e1;
                    lb = fmaf(lb, alpha, warp_sum(e0 + e1));
                    ob.x *= alpha; ob.y *= alpha; ob.z *= alpha; ob.w *= alpha;
                    mb = mn;
                } else { Ps[qb * ATP_KT + lane] = 0.f; Ps[qb * ATP_KT + lane + 32] = 0.f; }
            } else { Ps[qb * ATP_KT + lane] = 0.f; Ps[qb * ATP_KT + lane + 32] = 0.f; }
            __syncwarp();
            // ---- P.V for both queries: lane l accumulates dims 4l .. 4l+3 over the tile's keys
            if (lane < D4) {
                const float *pa = Ps + qa * ATP_KT, *pb = Ps + qb * ATP_KT;
                for (uint32_t j = 0; j < nk; j++) {   // masked keys carry weight 0
                    const float4 v = *reinterpret_cast<const float4 *>(Vs + j * HD + lane * 4);
                    const float wa = pa[j], wb = pb[j];
                    oa.x = fmaf(v.x, wa, oa.x); oa.y = fmaf(v.y, wa, oa.y); oa.z = fmaf(v.z, wa, oa.z); oa.w = fmaf(v.w, wa, oa.w);
                    ob.x = fmaf(v.x, wb, ob.x); ob.y = fmaf(v.y, wb, ob.y); ob.z = fmaf(v.z, wb, ob.z); ob.w = fmaf(v.w, wb, ob.w);
                }
            }
            __syncwarp();
        }
    }
    if (lane < D4) {
        if (qa < nq) {
            const float inv = __fdiv_rn(1.0f, la);
            *reinterpret_cast<float4 *>(out + (size_t)(q0 + qa) * dim + (size_t)h * HD + lane * 4) =
                make_float4(__fmul_rn(oa.x, inv), __fmul_rn(oa.y, inv), __fmul_rn(oa.z, inv), __fmul_rn(oa.w, inv));
        }
        if (qb < nq) {
            const float inv = __fdiv_rn(1.0f, lb);
            *reinterpret_cast<float4 *>(out + (size_t)(q0 + qb) * dim + (size_t)h * HD + lane * 4) =
                make_float4(__fmul_rn(ob.x, inv), __fmul_rn(ob.y, inv), __fmul_rn(ob.z, inv), __fmul_rn(ob.w, inv));
        }
    }
}

template <int HD>
static void attention_tiled_launch(const float *q, uint32_t ldq, const float *Kc, const float *Vc, float *out, uint32_t N,
                                   const uint32_t *past_dev, uint32_t dim, uint32_t heads, float scale, cudaStream_t st) {
    const size_t smem = ((size_t)ATP_KT * (HD + 4) + (size_t)ATP_KT * HD + (size_t)ATP_QT * HD + (size_t)ATP_QT * ATP_KT) * sizeof(float);
    static bool attr_set[64] = {};  // function attributes are per device
    int dev = 0;
    LB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        LB_CUDA(cudaFuncSetAttribute(attention_tiled_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    attention_tiled_kernel<HD><<<dim3(heads, (N + ATP_QT - 1) / ATP_QT), ATP_THREADS, smem, st>>>(q, ldq, Kc, Vc, out, past_dev, N, dim, scale);
    LB_LAUNCH_CHECK();
}

void attention(const float *q, uint32_t ldq, const float *Kc, const float *Vc, float *out, uint32_t N,
               const uint32_t *past_dev, uint32_t max_T, uint32_t dim, uint32_t heads, cudaStream_t st) {
    const uint32_t hd = dim / heads;
    LB_CHECK(hd == 32 || hd == 64 || hd == 128, "attention: head dim must be 32, 64 or 128");
    static const bool single_pass = getenv("LB_ATTN_SINGLE_PASS") != nullptr;   // A/B switch: the round-1 kernel
    if (!single_pass) {
        const float sc = (float)(1.0 / sqrt((double)dim / (double)heads));  // llama.go:306
        if (hd == 128) attention_tiled_launch<128>(q, ldq, Kc, Vc, out, N, past_dev, dim, heads, sc, st);
        else if (hd == 64) attention_tiled_launch<64>(q, ldq, Kc, Vc, out, N, past_dev, dim, heads, sc, st);
        else attention_tiled_launch<32>(q, ldq, Kc, Vc, out, N, past_dev, dim, heads, sc, st);
        return;
    }
    const uint32_t T = max_T;
    size_t smem = (((size_t)T + 3) & ~(size_t)3) * sizeof(float) + (size_t)ATT_THREADS * sizeof(float);
    static bool attr_set[64] = {};  // function attributes are per device
    int dev = 0;
    LB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        LB_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    LB_CHECK(smem <= 200 * 1024, "attention: context too long for the single-pass kernel");
    float scale = (float)(1.0 / sqrt((double)dim / (double)heads));  // llama.go:306
    attention_kernel<<<dim3(heads, N), ATT_THREADS, smem, st>>>(q, ldq, Kc, Vc, out, past_dev, max_T, dim, hd, scale);
    LB_LAUNCH_CHECK();
}

}  // namespace k
}  // namespace lb
