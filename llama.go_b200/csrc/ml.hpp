// ml.hpp — host-side mirror of the reference's pkg/ml tensor/op API on device memory.
//
// Same names, argument meaning and error behaviour as pkg/ml/ml.go (cited per function), so that
// pkg/llama.Eval's graph construction ports one to one; the data lives in HBM and GraphCompute
// walks the node list launching one CUDA kernel per op on the context's stream instead of
// fanning MulMat out to a goroutine pool (ml.go:1389-1399, 1602-1652).
// Where the reference prints "[HALT] ..." and os.Exit(1)s, these functions throw lb::Error with
// the same message; the C-ABI turns that into a non-zero status + lb_last_error().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <unordered_set>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace lb {
namespace ml {

constexpr uint32_t MAX_DIMS = 4;     // ml.go:19
constexpr uint32_t MAX_NODES = 4096; // ml.go:20

enum DType : int {  // ml.go:85-94
    TYPE_F32 = 0, TYPE_F16 = 1, TYPE_Q4_0 = 2, TYPE_Q4_1 = 3, TYPE_I8 = 4, TYPE_I16 = 5, TYPE_I32 = 6, TYPE_COUNT = 8
};

enum OpType : int {  // ml.go:133-174 (only the ops the reference implements are computable)
    OP_NONE = 0, OP_DUP, OP_ADD, OP_MUL, OP_REPEAT, OP_SILU, OP_RMS_NORM, OP_MUL_MAT, OP_SCALE, OP_CPY,
    OP_RESHAPE, OP_VIEW, OP_PERMUTE, OP_TRANSPOSE, OP_GET_ROWS, OP_DIAG_MASK_INF, OP_SOFT_MAX, OP_ROPE
};

struct Context;

// ml.Tensor (ml.go:180-203).  Data is a device pointer; `avail` = floats addressable from it
// (the Go slice's len), used to bound View1D/Reshape like Go's slice bounds checks would.
struct Tensor {
    DType type = TYPE_F32;
    uint32_t dims = 1;
    uint32_t ne[MAX_DIMS] = {1, 1, 1, 1};
    uint32_t nb[MAX_DIMS] = {4, 4, 4, 4};  // strides in BYTES, always 4-byte elements (ml.go:779)
    OpType op = OP_NONE;
    Tensor *src0 = nullptr, *src1 = nullptr;
    float *data = nullptr;
    size_t avail = 0;
    // host shadows of scalars the reference smuggles through tensor Data (ml.go:864-867, 927, 981)
    bool has_host = false;
    float host[3] = {0, 0, 0};

    uint32_t nelements() const { return ne[0] * ne[1] * ne[2] * ne[3]; }  // ml.go:217
    uint32_t nrows() const { return ne[1] * ne[2] * ne[3]; }              // ml.go:221
    bool is_contiguous() const {                                           // ml.go:206-211
        return nb[0] == 4 && nb[1] == nb[0] * ne[0] && nb[2] == nb[1] * ne[1] && nb[3] == nb[2] * ne[2];
    }
    TView view() const {
        TView v;
        v.data = data;
        for (int i = 0; i < 4; i++) { v.ne[i] = ne[i]; v.nb[i] = nb[i] / 4; }
        return v;
    }
};

// ml.Context (ml.go:50-57): owns the stream and every tensor/buffer created through it.
struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    std::vector<std::unique_ptr<Tensor>> tensors;
    std::vector<void *> buffers;

    explicit Context(int dev, cudaStream_t st = nullptr);
    ~Context();
    float *alloc(size_t floats);
    Tensor *track(std::unique_ptr<Tensor> t);
};

struct Graph {  // ml.go:31-45
    std::vector<Tensor *> nodes, leafs;
    std::unordered_set<const Tensor *> seen;
};

// ---- constructors (same signatures as the Go API) ----
Tensor *NewTensor(Context *ctx, DType dt, uint32_t dims, uint32_t ne0, uint32_t ne1, uint32_t ne2, uint32_t ne3,
                  float *data, size_t avail);                                       // ml.go:760
Tensor *NewTensor1D(Context *ctx, DType dt, uint32_t ne0);                          // ml.go:742
Tensor *NewTensor2D(Context *ctx, DType dt, uint32_t ne0, uint32_t ne1);            // ml.go:747
Tensor *NewTensor3D(Context *ctx, DType dt, uint32_t ne0, uint32_t ne1, uint32_t ne2);  // ml.go:751
Tensor *ViewTensor(Context *ctx, Tensor *src);                                      // ml.go:231
Tensor *DupTensor(Context *ctx, Tensor *src);                                       // ml.go:236
Tensor *NewFP32(Context *ctx, float value);                                         // ml.go:915

Tensor *GetRows(Context *ctx, Tensor *a, Tensor *b);                                // ml.go:528
Tensor *RMSNorm(Context *ctx, Tensor *a);                                           // ml.go:559
Tensor *Repeat(Context *ctx, Tensor *a, Tensor *b);                                 // ml.go:487
Tensor *Mul(Context *ctx, Tensor *a, Tensor *b);                                    // ml.go:241
Tensor *Add(Context *ctx, Tensor *a, Tensor *b);                                    // ml.go:347
Tensor *MulMat(Context *ctx, Tensor *a, Tensor *b);                                 // ml.go:295
Tensor *View1D(Context *ctx, Tensor *a, uint32_t ne0, uint32_t offset_floats);      // ml.go:601
Tensor *Copy(Context *ctx, Tensor *a, Tensor *b);                                   // ml.go:733
Tensor *Rope(Context *ctx, Tensor *a, uint32_t past, uint32_t dims, uint32_t mode); // ml.go:848
Tensor *Permute(Context *ctx, Tensor *a, uint32_t ax0, uint32_t ax1, uint32_t ax2, uint32_t ax3);  // ml.go:786
Tensor *Transpose(Context *ctx, Tensor *a);                                         // ml.go:1087
Tensor *Reshape3D(Context *ctx, Tensor *a, uint32_t ne0, uint32_t ne1, uint32_t ne2);  // ml.go:882
Tensor *Scale(Context *ctx, Tensor *a, Tensor *b);                                  // ml.go:959
Tensor *DiagMaskInf(Context *ctx, Tensor *a, uint32_t past);                        // ml.go:968
Tensor *SoftMax(Context *ctx, Tensor *a);                                           // ml.go:993
Tensor *Silu(Context *ctx, Tensor *a);                                              // ml.go:1041

void BuildForwardExpand(Graph *g, Tensor *t);    // ml.go:642
// ml.go:1411.  Enqueues every node on ctx->stream; synchronous like the reference unless sync=false.
void GraphCompute(Context *ctx, Graph *g, bool sync = true);

}  // namespace ml
}  // namespace lb
