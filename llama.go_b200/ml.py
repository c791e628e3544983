"""Python mirror of the reference's pkg/ml API (pkg/ml/ml.go), driving the C-ABI.

Same names and argument meaning as the Go package so parity tests read like Go call sites:
    ctx = ml.NewContext()
    a = ml.NewTensor2D(ctx, ml.TYPE_F32, ne0, ne1); a.set(host)
    c = ml.MulMat(ctx, w, x)
    g = ml.Graph(); ml.BuildForwardExpand(g, c); ml.GraphCompute(ctx, g)
    c.numpy()
Tensor data lives in HBM; `.Data` access of the Go struct becomes `.set()` / `.numpy()`.
Errors the reference reports with "[HALT] ..." + os.Exit(1) raise LlamaB200Error here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import LlamaB200Error, check, check_ptr, lib  # noqa: F401

TYPE_F32, TYPE_F16, TYPE_Q4_0, TYPE_Q4_1, TYPE_I8, TYPE_I16, TYPE_I32 = 0, 1, 2, 3, 4, 5, 6  # ml.go:85-94
_f32p = C.POINTER(C.c_float)


class Context:
    """ml.Context (ml.go:50-74).  maxThreads/useAVX/useNEON are accepted and ignored: the GPU
    replaces the goroutine pool."""

    def __init__(self, maxThreads: int = 0, useAVX: bool = False, useNEON: bool = False, device: int = 0):
        _capi.require_gpu()
        self._h = check_ptr(lib().lb_ml_new_context(device))
        self._keep = []

    def ReleaseContext(self):
        if self._h:
            lib().lb_ml_release_context(self._h)
            self._h = None

    def __del__(self):
        try:
            self.ReleaseContext()
        except Exception:
            pass


def NewContext(maxThreads: int = 0, useAVX: bool = False, useNEON: bool = False, device: int = 0) -> Context:
    return Context(maxThreads, useAVX, useNEON, device)


class Tensor:
    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = check_ptr(handle)

    @property
    def NE(self):
        ne, nb = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
        check(lib().lb_tensor_shape(self._h, ne, nb))
        return list(ne)

    @property
    def NB(self):
        ne, nb = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
        check(lib().lb_tensor_shape(self._h, ne, nb))
        return list(nb)

    def Nelements(self) -> int:
        return int(np.prod(self.NE))

    def set(self, host) -> "Tensor":
        a = np.ascontiguousarray(host, dtype=np.float32).reshape(-1)
        check(lib().lb_tensor_write(self._h, a.ctypes.data_as(_f32p), a.size))
        return self

    def numpy(self, nelem: int | None = None) -> np.ndarray:
        """The tensor's backing store as a flat array of Nelements (contiguous tensors) or `nelem` floats."""
        n = self.Nelements() if nelem is None else nelem
        out = np.empty(n, np.float32)
        check(lib().lb_tensor_read(self._h, out.ctypes.data_as(_f32p), n))
        if nelem is None:
            ne = self.NE
            return out.reshape(ne[3], ne[2], ne[1], ne[0])
        return out


def NewTensor(ctx, dt, dims, ne0, ne1=1, ne2=1, ne3=1, data=None) -> Tensor:
    host = None
    if data is not None:
        a = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
        assert a.size == ne0 * ne1 * ne2 * ne3
        host = a.ctypes.data_as(_f32p)
    return Tensor(ctx, lib().lb_new_tensor(ctx._h, dt, dims, ne0, ne1, ne2, ne3, host))


def NewTensor1D(ctx, dt, ne0, data=None):
    return NewTensor(ctx, dt, 1, ne0, 1, 1, 1, data)


def NewTensor2D(ctx, dt, ne0, ne1, data=None):
    return NewTensor(ctx, dt, 2, ne0, ne1, 1, 1, data)


def NewTensor3D(ctx, dt, ne0, ne1, ne2, data=None):
    return NewTensor(ctx, dt, 3, ne0, ne1, ne2, 1, data)


def NewTensor4D(ctx, dt, ne0, ne1, ne2, ne3, data=None):
    return NewTensor(ctx, dt, 4, ne0, ne1, ne2, ne3, data)


def _t(ctx, h):
    return Tensor(ctx, h)


def GetRows(ctx, a, b): return _t(ctx, lib().lb_get_rows(ctx._h, a._h, b._h))
def RMSNorm(ctx, a): return _t(ctx, lib().lb_rms_norm(ctx._h, a._h))
def Repeat(ctx, a, b): return _t(ctx, lib().lb_repeat(ctx._h, a._h, b._h))
def Mul(ctx, a, b): return _t(ctx, lib().lb_mul(ctx._h, a._h, b._h))
def Add(ctx, a, b): return _t(ctx, lib().lb_add(ctx._h, a._h, b._h))
def MulMat(ctx, a, b): return _t(ctx, lib().lb_mul_mat(ctx._h, a._h, b._h))
def View1D(ctx, a, ne0, offset): return _t(ctx, lib().lb_view_1d(ctx._h, a._h, ne0, offset))
def Copy(ctx, a, b): return _t(ctx, lib().lb_cpy(ctx._h, a._h, b._h))
def Rope(ctx, a, past, dims, mode): return _t(ctx, lib().lb_rope(ctx._h, a._h, past, dims, mode))
def Permute(ctx, a, ax0, ax1, ax2, ax3): return _t(ctx, lib().lb_permute(ctx._h, a._h, ax0, ax1, ax2, ax3))
def Transpose(ctx, a): return _t(ctx, lib().lb_transpose(ctx._h, a._h))
def Reshape3D(ctx, a, ne0, ne1, ne2): return _t(ctx, lib().lb_reshape_3d(ctx._h, a._h, ne0, ne1, ne2))
def NewFP32(ctx, value): return _t(ctx, lib().lb_new_f32(ctx._h, float(value)))
def Scale(ctx, a, b): return _t(ctx, lib().lb_scale(ctx._h, a._h, b._h))
def DiagMaskInf(ctx, a, past): return _t(ctx, lib().lb_diag_mask_inf(ctx._h, a._h, past))
def SoftMax(ctx, a): return _t(ctx, lib().lb_soft_max(ctx._h, a._h))
def Silu(ctx, a): return _t(ctx, lib().lb_silu(ctx._h, a._h))


class Vocab:
    """ml.Vocab (ml.go:2653-2657): tokens as bytes, with scores.  Host only (no GPU needed)."""

    def __init__(self, tokens, scores=None):
        self.tokens = [bytes(t) for t in tokens]
        self._h = check_ptr(lib().lb_vocab_create(len(self.tokens)))
        for i, t in enumerate(self.tokens):
            check(lib().lb_vocab_set(self._h, i, t, len(t), float(scores[i]) if scores is not None else 0.0))

    def __del__(self):
        try:
            lib().lb_vocab_free(self._h)
        except Exception:
            pass


def Tokenize(vocab: Vocab, text, bos: bool = True):
    """ml.Tokenize (ml.go:2761-2848)."""
    raw = text.encode() if isinstance(text, str) else bytes(text)
    cap = len(raw) + 2
    out = (C.c_uint32 * cap)()
    n = lib().lb_tokenize(vocab._h, raw, len(raw), 1 if bos else 0, out, cap)
    if n < 0:
        raise LlamaB200Error("lb_tokenize: bad arguments")
    return list(out[:n])


class Graph:
    """ml.Graph (ml.go:31-45)."""

    def __init__(self):
        self._h = check_ptr(lib().lb_graph_new())

    @property
    def NodesCount(self) -> int:
        return lib().lb_graph_nodes(self._h)

    def __del__(self):
        try:
            lib().lb_graph_free(self._h)
        except Exception:
            pass


def BuildForwardExpand(graph: Graph, tensor: Tensor) -> None:
    check(lib().lb_build_forward_expand(graph._h, tensor._h))


def GraphCompute(ctx: Context, graph: Graph) -> None:
    check(lib().lb_graph_compute(ctx._h, graph._h))
