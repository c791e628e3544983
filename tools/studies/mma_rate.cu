// Throughput of the legacy warp-level tensor instructions on this part (sm_100a): how many
// mma.sync.m16n8k8.tf32 / m16n8k32.s8 / m16n8k16.bf16 per clock and SM.  Each warp runs NACC independent
// accumulator chains; 148 x 4 CTAs of 8 warps.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int NACC>
__global__ void k(float *out, int iters) {
    float d[NACC][4];
    int di[NACC][4];
    for (int a = 0; a < NACC; a++) for (int i = 0; i < 4; i++) { d[a][i] = 0.f; di[a][i] = 0; }
    unsigned a0 = threadIdx.x * 0x3f800001u, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 ^ 0x5555, b1 = b0 + 7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int a = 0; a < NACC; a++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[a][0]), "+f"(d[a][1]), "+f"(d[a][2]), "+f"(d[a][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(di[a][0]), "+r"(di[a][1]), "+r"(di[a][2]), "+r"(di[a][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[a][0]), "+f"(d[a][1]), "+f"(d[a][2]), "+f"(d[a][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0;
    for (int a = 0; a < NACC; a++) for (int i = 0; i < 4; i++) s += d[a][i] + (float)di[a][i];
    if (s == 123.456f) out[0] = s;
}
template <int KIND, int NACC>
void run(const char *name, double macs_per_mma) {
    float *out; cudaMalloc(&out, 4);
    const int iters = 20000, ctas = 148 * 4, thr = 256;
    k<KIND, NACC><<<ctas, thr>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND, NACC><<<ctas, thr>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = (double)ctas * (thr / 32) * iters * NACC;
    double per_sm_clk = mmas / 148.0 / (ms * 1e-3 * 1.965e9);
    printf("%-28s NACC=%d: %.3f ms, %.3f mma/clk/SM (at 1965 MHz), %.0f MAC/clk/SM, %.1f T MAC/s\n", name, NACC, ms, per_sm_clk,
           per_sm_clk * macs_per_mma, mmas * macs_per_mma / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    run<0, 1>("m16n8k8 tf32 (dependent)", 1024); run<0, 4>("m16n8k8 tf32", 1024); run<0, 8>("m16n8k8 tf32", 1024);
    run<1, 4>("m16n8k32 s8", 4096); run<1, 8>("m16n8k32 s8", 4096);
    run<2, 4>("m16n8k16 bf16", 2048); run<2, 8>("m16n8k16 bf16", 2048);
    return 0;
}
