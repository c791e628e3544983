// How fast does one SM pull HBM data through cp.async.bulk (1-D bulk copies, mbarrier completion) as a function of
// the copy size?  148 CTAs, each streams its own 64 MB slice of a 9.5 GB buffer through a ring of 8 x 16 KB slots:
// a producer warp fills a slot with 16 KB / S copies of S bytes (issued by `L` lanes in parallel), one consumer warp
// waits for the slot and releases it at once.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tma_copy_rate.cu -o tma_copy_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
constexpr int NS = 8, SLOT = 16384;
__global__ void __launch_bounds__(64) k(const uint8_t *src, size_t per_cta, int S, int L, unsigned long long *sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ unsigned long long full[NS], empty[NS];
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[i])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[i])), "r"(1));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint8_t *base = src + (size_t)blockIdx.x * per_cta;
    const size_t nslots = per_cta / SLOT;
    const int per_slot = SLOT / S;
    if (threadIdx.x < 32) {
        for (size_t q = 0; q < nslots; q++) {
            const int s = q % NS;
            const uint32_t ph = (q / NS) & 1;
            if (lane == 0) {
                mbar_wait(s32(&empty[s]), ph ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(SLOT) : "memory");
            }
            __syncwarp();
            for (int c = lane; c < per_slot; c += L) {
                if (lane < L)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(s32(sm + s * SLOT + c * S)), "l"(base + q * SLOT + (size_t)c * S), "r"(S), "r"(s32(&full[s])) : "memory");
            }
        }
    } else {
        unsigned long long acc = 0;
        for (size_t q = 0; q < nslots; q++) {
            const int s = q % NS;
            const uint32_t ph = (q / NS) & 1;
            mbar_wait(s32(&full[s]), ph);
            acc += *reinterpret_cast<const unsigned long long *>(sm + s * SLOT + lane * 8);
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        }
        if (acc == 0x1234567) sink[0] = acc;
    }
}
int main() {
    const size_t per_cta = 64ull << 20, total = per_cta * 148;
    uint8_t *src; unsigned long long *sink;
    cudaMalloc(&src, total); cudaMalloc(&sink, 8); cudaMemset(src, 1, total);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, NS * SLOT);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int sizes[] = {512, 1024, 2048, 4096, 8192, 16384};
    for (int S : sizes)
        for (int L : {1, 16, 32}) {
            if (SLOT / S < L && L > 1 && SLOT / S < 16) { if (L != 1 && SLOT / S < L && L == 32) continue; }
            k<<<148, 64, NS * SLOT>>>(src, per_cta, S, L, sink);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k<<<148, 64, NS * SLOT>>>(src, per_cta, S, L, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("copy %5d B, %2d issuing lanes: %.3f ms  %.0f GB/s  (%.1f B/clk/SM at 1965 MHz)  err=%s\n", S, L, ms, total / (ms * 1e-3) / 1e9,
                   total / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
