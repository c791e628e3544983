// common.cuh — shared host/device helpers for the llamab200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <stdexcept>
#include <string>
#include <vector>

namespace lb {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define LB_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            throw lb::Error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                            ":" + std::to_string(__LINE__) + ")");                             \
    } while (0)

#define LB_CHECK(cond, msg)                                    \
    do {                                                       \
        if (!(cond)) throw lb::Error(std::string("[HALT] ") + (msg)); \
    } while (0)

// process-wide count of kernels this library launched (bench.py's gpu_launches)
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define LB_LAUNCH_CHECK()            \
    do {                             \
        lb::count_launch();          \
        LB_CUDA(cudaGetLastError()); \
    } while (0)

constexpr int kNumSMs = 148;

// Owns device / pinned-host allocations, events and streams of one object and releases them in its
// destructor — declared as the FIRST member, so an exception half-way through a constructor (e.g. an
// out-of-memory cudaMalloc) does not leak what was already allocated.
struct DeviceOwner {
    int device = 0;
    std::vector<void *> dev, pinned;
    std::vector<cudaEvent_t> events;
    std::vector<cudaStream_t> streams;
    DeviceOwner() = default;
    DeviceOwner(const DeviceOwner &) = delete;
    DeviceOwner &operator=(const DeviceOwner &) = delete;
    ~DeviceOwner() {
        cudaSetDevice(device);
        for (cudaStream_t s : streams) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        for (void *p : dev) cudaFree(p);
        for (void *p : pinned) cudaFreeHost(p);
    }
    template <typename T>
    T *dmalloc(size_t n, bool zero = true) {
        void *p = nullptr;
        const size_t bytes = (n ? n : 1) * sizeof(T);
        LB_CUDA(cudaMalloc(&p, bytes));
        dev.push_back(p);
        if (zero) LB_CUDA(cudaMemset(p, 0, bytes));
        return static_cast<T *>(p);
    }
    template <typename T>
    T *hmalloc(size_t n) {
        void *p = nullptr;
        LB_CUDA(cudaMallocHost(&p, (n ? n : 1) * sizeof(T)));
        pinned.push_back(p);
        return static_cast<T *>(p);
    }
    cudaEvent_t event() {
        cudaEvent_t e = nullptr;
        LB_CUDA(cudaEventCreate(&e));
        events.push_back(e);
        return e;
    }
    cudaStream_t stream() {
        cudaStream_t s = nullptr;
        LB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        streams.push_back(s);
        return s;
    }
};

// Programmatic dependent launch (PDL): decode kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that kernel k+1's CTAs are scheduled into the
// tail of kernel k; each such kernel issues its weight prefetch first, then `griddepcontrol.wait`
// (= cudaGridDependencySynchronize) before it touches anything the predecessor wrote.
extern bool g_use_pdl;
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_use_pdl ? 1 : 0;
    LB_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
    count_launch();
}

#ifdef __CUDACC__
// PDL device side: let the dependent grid start launching / wait for the predecessor grid's memory
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// streaming 128-bit load: read-only path, do not allocate in L1 (weights are touched once)
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ int4 ld_stream_i4(const void *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// SiluFP32 of the reference: x / float32(1 + exp(float64(-x)))   (pkg/ml/ml.go:2587-2589)
__device__ __forceinline__ float silu_ref(float x) {
    return __fdiv_rn(x, (float)(1.0 + exp((double)(-x))));
}
#endif

}  // namespace lb
