// ml.cpp — see ml.hpp.  Compiled by nvcc (host code only).
#include "ml.hpp"

#include <string.h>

namespace lb {
namespace ml {

Context::Context(int dev, cudaStream_t st) : device(dev) {
    LB_CUDA(cudaSetDevice(dev));
    if (st) {
        stream = st;
    } else {
        LB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        owns_stream = true;
    }
}
Context::~Context() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (void *p : buffers) cudaFree(p);
    if (owns_stream && stream) cudaStreamDestroy(stream);
}
float *Context::alloc(size_t floats) {
    LB_CUDA(cudaSetDevice(device));
    void *p = nullptr;
    size_t bytes = (floats ? floats : 1) * sizeof(float);
    LB_CUDA(cudaMalloc(&p, bytes));
    LB_CUDA(cudaMemsetAsync(p, 0, bytes, stream));  // Go's make() zero-fills (ml.go:770-773)
    buffers.push_back(p);
    return static_cast<float *>(p);
}
Tensor *Context::track(std::unique_ptr<Tensor> t) {
    tensors.push_back(std::move(t));
    return tensors.back().get();
}

Tensor *NewTensor(Context *ctx, DType dt, uint32_t dims, uint32_t ne0, uint32_t ne1, uint32_t ne2, uint32_t ne3,
                  float *data, size_t avail) {
    LB_CHECK(ctx != nullptr, "NewTensor : nil context");
    auto t = std::make_unique<Tensor>();
    t->type = dt;
    t->dims = dims;
    t->ne[0] = ne0; t->ne[1] = ne1; t->ne[2] = ne2; t->ne[3] = ne3;
    t->nb[0] = 4; t->nb[1] = ne0 * 4; t->nb[2] = ne0 * ne1 * 4; t->nb[3] = ne0 * ne1 * ne2 * 4;
    size_t total = (size_t)ne0 * ne1 * ne2 * ne3;
    if (data == nullptr) {
        t->data = ctx->alloc(total);
        t->avail = total;
    } else {
        t->data = data;
        t->avail = avail;
    }
    return ctx->track(std::move(t));
}
Tensor *NewTensor1D(Context *ctx, DType dt, uint32_t ne0) { return NewTensor(ctx, dt, 1, ne0, 1, 1, 1, nullptr, 0); }
Tensor *NewTensor2D(Context *ctx, DType dt, uint32_t ne0, uint32_t ne1) { return NewTensor(ctx, dt, 2, ne0, ne1, 1, 1, nullptr, 0); }
Tensor *NewTensor3D(Context *ctx, DType dt, uint32_t ne0, uint32_t ne1, uint32_t ne2) {
    return NewTensor(ctx, dt, 3, ne0, ne1, ne2, 1, nullptr, 0);
}
Tensor *ViewTensor(Context *ctx, Tensor *src) {
    // NB: like the reference this rebuilds contiguous strides from NE (ml.go:231-233)
    return NewTensor(ctx, src->type, src->dims, src->ne[0], src->ne[1], src->ne[2], src->ne[3], src->data, src->avail);
}
Tensor *DupTensor(Context *ctx, Tensor *src) {
    return NewTensor(ctx, src->type, src->dims, src->ne[0], src->ne[1], src->ne[2], src->ne[3], nullptr, 0);
}
Tensor *NewFP32(Context *ctx, float value) {
    Tensor *t = NewTensor1D(ctx, TYPE_F32, 1);
    LB_CUDA(cudaMemcpyAsync(t->data, &value, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    LB_CUDA(cudaStreamSynchronize(ctx->stream));  // `value` is a stack temporary
    t->has_host = true;
    t->host[0] = value;
    return t;
}

static bool same_shape(const Tensor *a, const Tensor *b) {  // ml.go:213
    return a->ne[0] == b->ne[0] && a->ne[1] == b->ne[1] && a->ne[2] == b->ne[2] && a->ne[3] == b->ne[3];
}

Tensor *GetRows(Context *ctx, Tensor *a, Tensor *b) {
    Tensor *r = NewTensor2D(ctx, TYPE_F32, a->ne[0], b->ne[0]);
    r->op = OP_GET_ROWS; r->src0 = a; r->src1 = b;
    return r;
}
Tensor *RMSNorm(Context *ctx, Tensor *a) {
    Tensor *r = DupTensor(ctx, a);
    r->op = OP_RMS_NORM; r->src0 = a;
    return r;
}
Tensor *Repeat(Context *ctx, Tensor *a, Tensor *b) {
    if (same_shape(a, b)) return a;  // ml.go:496-498
    Tensor *r = NewTensor(ctx, a->type, b->dims, b->ne[0], b->ne[1], b->ne[2], b->ne[3], nullptr, 0);
    r->op = OP_REPEAT; r->src0 = a; r->src1 = b;
    return r;
}
Tensor *Mul(Context *ctx, Tensor *a, Tensor *b) {
    LB_CHECK(same_shape(a, b), "MulImpl - tensors of different shapes!");  // ml.go:254-257
    Tensor *r = DupTensor(ctx, a);
    r->op = OP_MUL; r->src0 = a; r->src1 = b;
    return r;
}
Tensor *Add(Context *ctx, Tensor *a, Tensor *b) {
    Tensor *r = DupTensor(ctx, a);
    r->op = OP_ADD; r->src0 = a; r->src1 = b;
    return r;
}
Tensor *MulMat(Context *ctx, Tensor *a, Tensor *b) {
    Tensor *r = NewTensor(ctx, TYPE_F32, a->dims < b->dims ? a->dims : b->dims, a->ne[1], b->ne[1], a->ne[2], b->ne[3],
                          nullptr, 0);  // ml.go:305
    r->op = OP_MUL_MAT; r->src0 = a; r->src1 = b;
    return r;
}
Tensor *View1D(Context *ctx, Tensor *a, uint32_t ne0, uint32_t offset) {
    LB_CHECK((size_t)offset + ne0 <= a->avail, "View1D : offset + ne0 out of range");  // Go: slice bounds panic
    Tensor *r = NewTensor(ctx, a->type, 1, ne0, 1, 1, 1, a->data + offset, a->avail - offset);
    r->op = OP_VIEW; r->src0 = a;
    return r;
}
Tensor *Copy(Context *ctx, Tensor *a, Tensor *b) {
    Tensor *r = ViewTensor(ctx, b);  // ml.go:718
    r->op = OP_CPY; r->src0 = a; r->src1 = b;
    return r;
}
Tensor *Rope(Context *ctx, Tensor *a, uint32_t past, uint32_t dims, uint32_t mode) {
    Tensor *r = ViewTensor(ctx, a);  // in place, ml.go:862
    Tensor *b = NewTensor(ctx, TYPE_I32, 1, 3, 1, 1, 1, nullptr, 0);
    b->has_host = true;
    b->host[0] = (float)past; b->host[1] = (float)dims; b->host[2] = (float)mode;
    LB_CUDA(cudaMemcpyAsync(b->data, b->host, 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    r->op = OP_ROPE; r->src0 = a; r->src1 = b;
    return r;
}
Tensor *Permute(Context *ctx, Tensor *a, uint32_t ax0, uint32_t ax1, uint32_t ax2, uint32_t ax3) {
    LB_CHECK(ax0 < 4 && ax1 < 4 && ax2 < 4 && ax3 < 4 && ax0 != ax1 && ax0 != ax2 && ax0 != ax3 && ax1 != ax2 &&
                 ax1 != ax3 && ax2 != ax3, "Permute error");
    Tensor *r = ViewTensor(ctx, a);
    uint32_t ne[4], nb[4];
    ne[ax0] = a->ne[0]; ne[ax1] = a->ne[1]; ne[ax2] = a->ne[2]; ne[ax3] = a->ne[3];
    nb[ax0] = a->nb[0]; nb[ax1] = a->nb[1]; nb[ax2] = a->nb[2]; nb[ax3] = a->nb[3];
    for (int i = 0; i < 4; i++) { r->ne[i] = ne[i]; r->nb[i] = nb[i]; }
    r->op = OP_PERMUTE; r->src0 = a;
    return r;
}
Tensor *Transpose(Context *ctx, Tensor *a) {
    Tensor *r = ViewTensor(ctx, a);
    r->ne[0] = a->ne[1]; r->ne[1] = a->ne[0];
    r->nb[0] = a->nb[1]; r->nb[1] = a->nb[0];
    r->op = OP_TRANSPOSE; r->src0 = a;
    return r;
}
Tensor *Reshape3D(Context *ctx, Tensor *a, uint32_t ne0, uint32_t ne1, uint32_t ne2) {
    LB_CHECK((size_t)ne0 * ne1 * ne2 <= a->avail, "Reshape3D : different elements number!");
    Tensor *r = NewTensor(ctx, a->type, 3, ne0, ne1, ne2, 1, a->data, a->avail);
    r->op = OP_RESHAPE; r->src0 = a;
    return r;
}
Tensor *Scale(Context *ctx, Tensor *a, Tensor *b) {
    Tensor *r = ViewTensor(ctx, a);
    r->op = OP_SCALE; r->src0 = a; r->src1 = b;
    return r;
}
Tensor *DiagMaskInf(Context *ctx, Tensor *a, uint32_t past) {
    Tensor *r = ViewTensor(ctx, a);
    Tensor *b = NewFP32(ctx, (float)past);  // ml.go:981
    r->op = OP_DIAG_MASK_INF; r->src0 = a; r->src1 = b;
    return r;
}
Tensor *SoftMax(Context *ctx, Tensor *a) {
    Tensor *r = ViewTensor(ctx, a);
    r->op = OP_SOFT_MAX; r->src0 = a;
    return r;
}
Tensor *Silu(Context *ctx, Tensor *a) {
    Tensor *r = DupTensor(ctx, a);
    r->op = OP_SILU; r->src0 = a;
    return r;
}

// ggml_visit_parents, ml.go:647-697 (DFS post-order, src0 before src1).  The reference's visited
// check is a linear scan (O(n^2)); a hash set gives the same order.
static void VisitParents(Graph *g, Tensor *node) {
    if (g->seen.count(node)) return;
    if (node->src0) VisitParents(g, node->src0);
    if (node->src1) VisitParents(g, node->src1);
    if (g->seen.count(node)) return;
    g->seen.insert(node);
    if (node->op == OP_NONE) {
        g->leafs.push_back(node);
    } else {
        LB_CHECK(g->nodes.size() < MAX_NODES, "graph : too many nodes (MAX_NODES = 4096)");
        g->nodes.push_back(node);
    }
}
void BuildForwardExpand(Graph *g, Tensor *t) {
    size_t n0 = g->nodes.size();
    VisitParents(g, t);
    if (g->nodes.size() > n0)
        LB_CHECK(g->nodes.back() == t, "BuildForwardImpl : the last added node should always be starting point!");
}

static float scalar_of(Context *ctx, const Tensor *t) {
    if (t->has_host) return t->host[0];
    float v;
    LB_CUDA(cudaMemcpyAsync(&v, t->data, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    LB_CUDA(cudaStreamSynchronize(ctx->stream));
    return v;
}

// ComputeForward, ml.go:1532-1702
static void ComputeForward(Context *ctx, Tensor *t) {
    cudaStream_t st = ctx->stream;
    Tensor *s0 = t->src0, *s1 = t->src1;
    switch (t->op) {
        case OP_ADD:
            LB_CHECK(s1->nb[0] == 4, "ComputeForwardAddFP32 : [src1] is NOT contiguous!");
            LB_CHECK(s0->is_contiguous() && s1->is_contiguous() && t->is_contiguous(), "ComputeForwardAddFP32 : strided rows not supported");
            LB_CHECK(s0->nelements() == t->nelements() && s1->nelements() == t->nelements(), "ComputeForwardAddFP32 : different element counts!");
            LB_CHECK(t->nelements() <= t->avail && s0->nelements() <= s0->avail && s1->nelements() <= s1->avail, "ComputeForwardAddFP32 : tensor exceeds its storage");  // Go: slice bounds panic
            k::add(s0->data, s1->data, t->data, t->nelements(), st);
            break;
        case OP_MUL:
            LB_CHECK(same_shape(s0, s1) && same_shape(s0, t), "ComputeForwardMulFP32 : different shapes!");
            k::mul(s0->data, s1->data, t->data, t->nelements(), st);
            break;
        case OP_REPEAT:
            k::repeat_rows(s0->data, s0->ne[0], s0->ne[1], t->data, t->ne[0], t->ne[1], st);
            break;
        case OP_SILU:
            LB_CHECK(s0->is_contiguous(), "ComputeForwardSiluFP32 : [src0] is NOT contiguous!");
            LB_CHECK(t->is_contiguous(), "ComputeForwardSiluFP32 : [dst] is NOT contiguous!");
            k::silu(s0->data, t->data, t->nelements(), st);
            break;
        case OP_RMS_NORM:
            LB_CHECK(s0->is_contiguous() && t->is_contiguous(), "ComputeForwardRMSNormFP32 : strided rows not supported");
            k::rms_norm(s0->data, nullptr, t->data, s0->ne[0], s0->nrows(), st);
            break;
        case OP_MUL_MAT: {
            LB_CHECK(s0->ne[0] == s1->ne[0] && s0->ne[2] == s1->ne[2] && s0->ne[3] == s1->ne[3], "MulMat : incompatible shapes");
            LB_CHECK(s0->nb[0] == 4 && s1->nb[0] == 4, "MulMat : transposed operands are not supported");
            const bool plain2d = s0->is_contiguous() && s1->is_contiguous() && s0->ne[2] == 1 && s0->ne[3] == 1 &&
                                 s1->ne[2] == 1 && s1->ne[3] == 1 && (s0->ne[0] & 3) == 0;
            if (plain2d && s1->ne[1] <= 8)
                k::gemv_f32(s0->data, s0->ne[1], s0->ne[0], s1->data, s1->ne[0], s1->ne[1], t->data, t->ne[0], nullptr, st);
            else if (plain2d)
                k::gemm_auto(s0->data, s0->ne[1], s0->ne[0], s1->data, s1->ne[0], s1->ne[1], t->data, t->ne[0], nullptr, st);
            else
                k::mul_mat_generic(s0->view(), s1->view(), t->view(), st);
            break;
        }
        case OP_SCALE:
            LB_CHECK(s0->is_contiguous(), "ComputeForwardScaleFP32 : [src0] is NOT contiguous!");
            LB_CHECK(t->is_contiguous(), "ComputeForwardScaleFP32 : [dst] is NOT contiguous!");
            LB_CHECK(t->nelements() <= t->avail, "ComputeForwardScaleFP32 : tensor exceeds its storage");
            k::scale_inplace(t->data, scalar_of(ctx, s1), t->nelements(), st);
            break;
        case OP_CPY:
            LB_CHECK(t->is_contiguous(), "ComputeForwardDupFP32 : [dst] is NOT contiguous!");
            LB_CHECK(t->nelements() == s0->nelements(), "ComputeForwardDupFP32 : [dst] and [src0] capacities are different!");
            LB_CHECK(t->nelements() <= t->avail, "ComputeForwardDupFP32 : [dst] exceeds its storage");  // Go: slice bounds panic
            LB_CHECK(!s0->is_contiguous() || s0->nelements() <= s0->avail, "ComputeForwardDupFP32 : [src0] exceeds its storage");
            if (s0->is_contiguous())
                LB_CUDA(cudaMemcpyAsync(t->data, s0->data, (size_t)t->nelements() * 4, cudaMemcpyDeviceToDevice, st));
            else
                k::cpy_strided(s0->view(), t->data, st);
            break;
        case OP_RESHAPE: case OP_VIEW: case OP_PERMUTE:
            break;  // NOP (ml.go:2101, 2243, 2248)
        case OP_TRANSPOSE:
            LB_CHECK(false, "Please implement : ggml_compute_forward_transpose");
            break;
        case OP_GET_ROWS:
            LB_CHECK(t->ne[0] == s0->ne[0] && t->ne[1] == s1->nelements() && s0->nb[0] == 4, "ComputeForwardGetRows : wrong dimensions!");
            k::get_rows_f32ids(s0->data, s0->ne[0], s1->data, s1->nelements(), t->data, st);
            break;
        case OP_DIAG_MASK_INF:
            k::diag_mask_inf(t->data, s0->ne[0], s0->ne[1], s0->ne[2] * s0->ne[3], (uint32_t)scalar_of(ctx, s1), st);
            break;
        case OP_SOFT_MAX:
            LB_CHECK(s0->is_contiguous(), "ComputeForwardSoftMaxFP32 : [src0] is NOT contiguous!");
            LB_CHECK(t->is_contiguous(), "ComputeForwardSoftMaxFP32 : [dst] is NOT contiguous!");
            k::soft_max_rows(t->data, t->ne[0], t->nrows(), st);
            break;
        case OP_ROPE: {
            LB_CHECK(s1->nelements() == 3 && s1->has_host, "ComputeForwardRopeFP32 : src1 has NOT EXACT 3 elements!");
            LB_CHECK(s0->is_contiguous() && s0->ne[3] == 1, "ComputeForwardRopeFP32 : strided input not supported");
            k::rope(t->data, s0->ne[0], s0->ne[1], s0->ne[2], (uint32_t)s1->host[0], (uint32_t)s1->host[1], (uint32_t)s1->host[2], st);
            break;
        }
        case OP_NONE:
            break;
        default:
            LB_CHECK(false, "ComputeForward : unsupported op");
    }
}

void GraphCompute(Context *ctx, Graph *g, bool sync) {
    LB_CUDA(cudaSetDevice(ctx->device));
    for (Tensor *node : g->nodes) ComputeForward(ctx, node);
    if (sync) LB_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace ml
}  // namespace lb
