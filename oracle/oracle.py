"""ctypes wrapper around oracle/liboracle.so (llama_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of llama_oracle.c.  Imported by tests/, by
bench.py's cpu_baseline / --impl reference legs and by __graft_entry__.smoke(); never by
anything under llama.go_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)


def build(force: bool = False) -> str:
    """Compile liboracle.so (and, when /root/reference is present, oracle/_ref)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "llama_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/builds"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.lo_model_new.restype = C.c_void_p
        L.lo_model_new.argtypes = [C.c_uint32] * 5
        L.lo_model_free.argtypes = [C.c_void_p]
        L.lo_model_ff.restype = C.c_uint32
        L.lo_model_ff.argtypes = [C.c_void_p]
        L.lo_model_tensor.restype = _f32p
        L.lo_model_tensor.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64)]
        L.lo_context_new.restype = C.c_void_p
        L.lo_context_new.argtypes = [C.c_void_p, C.c_uint32]
        L.lo_context_free.argtypes = [C.c_void_p]
        L.lo_context_k.restype = _f32p
        L.lo_context_k.argtypes = [C.c_void_p]
        L.lo_context_v.restype = _f32p
        L.lo_context_v.argtypes = [C.c_void_p]
        L.lo_eval.restype = C.c_int
        L.lo_eval.argtypes = [C.c_void_p, _u32p, C.c_uint32, C.c_uint32, _f32p, _f32p, _f32p]
        L.lo_set_dot_mode.argtypes = [C.c_int]
        L.lo_set_threads.argtypes = [C.c_int]
        L.lo_vdot.restype = C.c_float
        L.lo_vdot.argtypes = [_f32p, _f32p, C.c_uint32, C.c_int]
        _LIB = L
    return _LIB


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def _c(a, dtype=np.float32):
    return np.ascontiguousarray(a, dtype=dtype)


class OracleModel:
    """Host weights in the oracle's own storage (filled by name, ggjt naming)."""

    def __init__(self, hp):
        self.hp = hp
        self._h = lib().lo_model_new(hp.vocab, hp.dim, hp.mult, hp.heads, hp.layers)
        assert lib().lo_model_ff(self._h) == hp.ff

    def set_tensor(self, name: str, arr: np.ndarray) -> None:
        n = C.c_uint64(0)
        p = lib().lo_model_tensor(self._h, name.encode(), C.byref(n))
        if not p or n.value != arr.size:
            raise KeyError(f"oracle: bad tensor {name} ({arr.size} vs {n.value})")
        dst = np.ctypeslib.as_array(p, shape=(n.value,))
        dst[:] = _c(arr).reshape(-1)

    def load(self, tensors) -> "OracleModel":
        for name, arr in tensors:
            self.set_tensor(name, arr)
        return self

    def __del__(self):
        try:
            lib().lo_model_free(self._h)
        except Exception:
            pass


class OracleContext:
    """Mirror of llama.Context (pkg/llama/llama.go:83-113): owns the FP32 KV cache."""

    def __init__(self, model: OracleModel, ctx_size: int):
        self.model = model
        self.ctx_size = ctx_size
        self._h = lib().lo_context_new(model._h, ctx_size)

    def eval(self, tokens, past: int, all_logits: bool = False, hidden: bool = False):
        hp = self.model.hp
        toks = _c(tokens, np.uint32)
        n = toks.size
        logits = np.empty(hp.vocab, np.float32)
        allb = np.empty((n, hp.vocab), np.float32) if all_logits else None
        hid = np.empty((n, hp.dim), np.float32) if hidden else None
        rc = lib().lo_eval(self._h, toks.ctypes.data_as(_u32p), n, past, _fp(logits),
                           _fp(allb) if allb is not None else None,
                           _fp(hid) if hid is not None else None)
        if rc != 0:
            raise ValueError(f"oracle lo_eval rc={rc}")
        out = [logits]
        if all_logits:
            out.append(allb)
        if hidden:
            out.append(hid)
        return out[0] if len(out) == 1 else tuple(out)

    def kv(self):
        hp = self.model.hp
        n = hp.layers * self.ctx_size * hp.dim
        k = np.ctypeslib.as_array(lib().lo_context_k(self._h), shape=(n,)).reshape(hp.layers, self.ctx_size, hp.dim)
        v = np.ctypeslib.as_array(lib().lo_context_v(self._h), shape=(n,)).reshape(hp.layers, self.ctx_size, hp.dim)
        return k, v

    def __del__(self):
        try:
            lib().lo_context_free(self._h)
        except Exception:
            pass


def synth_fill(seed: int, tid: int, start: int, count: int, mean: float, sigma: float) -> np.ndarray:
    """Elements [start, start+count) of synthetic tensor `tid` — the recipe of llama.go_b200/synth.py
    synth_values(), multi-threaded in liboracle.so.  Lets bench.py's reference arm write the 27 GB ggjt file
    without loading the product library."""
    f = lib().lo_synth_fill
    f.restype = None
    f.argtypes = [_f32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, C.c_double]
    out = np.empty(count, np.float32)
    f(_fp(out), count, seed, tid, start, float(mean), float(sigma))
    return out


def synth_model(seed: int, hp, tensor_table):
    """(name, ndarray) for every row of synth.tensor_table(hp), generated by synth_fill()."""
    for name, tid, shape, mean, sigma in tensor_table:
        yield name, synth_fill(seed, tid, 0, int(np.prod(shape)), mean, sigma).reshape(shape)


def set_dot_mode(avx: bool) -> None:
    lib().lo_set_dot_mode(1 if avx else 0)


def set_threads(n: int) -> None:
    lib().lo_set_threads(n)


# ----------------------------------------------------------------------------- op-level wrappers
def _decl(name, argtypes):
    f = getattr(lib(), name)
    f.argtypes = argtypes
    f.restype = None
    return f


def _u4(v):
    return (C.c_uint32 * 4)(*[int(x) for x in v])


def op_get_rows(table, ids):
    table = _c(table)
    idsf = _c(ids, np.float32)
    out = np.empty((idsf.size, table.shape[1]), np.float32)
    _decl("lo_op_get_rows", [_f32p, C.c_uint32, _f32p, C.c_uint32, _f32p])(_fp(table), table.shape[1], _fp(idsf), idsf.size, _fp(out))
    return out


def op_rms_norm(x):
    x = _c(x)
    x2 = x.reshape(-1, x.shape[-1])
    y = np.empty_like(x2)
    _decl("lo_op_rms_norm", [_f32p, C.c_uint32, C.c_uint32, _f32p])(_fp(x2), x2.shape[1], x2.shape[0], _fp(y))
    return y.reshape(x.shape)


def op_repeat(a, rows):
    a = _c(a).reshape(1, -1)
    out = np.empty((rows, a.shape[1]), np.float32)
    _decl("lo_op_repeat", [_f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _f32p])(_fp(a), a.shape[1], 1, a.shape[1], rows, _fp(out))
    return out


def op_mul(a, b):
    a, b = _c(a), _c(b)
    out = np.empty_like(a)
    _decl("lo_op_mul", [_f32p, _f32p, C.c_uint32, _f32p])(_fp(a), _fp(b), a.size, _fp(out))
    return out


def op_add(a, b):
    a, b = _c(a), _c(b)
    out = np.empty_like(a)
    _decl("lo_op_add", [_f32p, _f32p, C.c_uint32, _f32p])(_fp(a), _fp(b), a.size, _fp(out))
    return out


def op_mul_mat(a, ne0, nb0, b, ne1, nb1):
    """General strided MulMat (strides in floats). Returns [ne1[3], ne0[2], ne1[1], ne0[1]] numpy (C order)."""
    a, b = _c(a), _c(b)
    out = np.empty((ne1[3], ne0[2], ne1[1], ne0[1]), np.float32)
    _decl("lo_op_mul_mat", [_f32p, C.c_uint32 * 4, C.c_uint32 * 4, _f32p, C.c_uint32 * 4, C.c_uint32 * 4, _f32p])(
        _fp(a), _u4(ne0), _u4(nb0), _fp(b), _u4(ne1), _u4(nb1), _fp(out))
    return out


def op_mul_mat_2d(w, x):
    """w [M,K] row-major, x [N,K] -> [N,M] (the weight MulMat of llama.go:263)."""
    w, x = _c(w), _c(x)
    M, K = w.shape
    N = x.shape[0]
    return op_mul_mat(w, (K, M, 1, 1), (1, K, K * M, K * M), x, (K, N, 1, 1), (1, K, K * N, K * N)).reshape(N, M)


def op_cpy(a, ne, nb):
    a = _c(a)
    out = np.empty((ne[3], ne[2], ne[1], ne[0]), np.float32)
    _decl("lo_op_cpy", [_f32p, C.c_uint32 * 4, C.c_uint32 * 4, _f32p])(_fp(a), _u4(ne), _u4(nb), _fp(out))
    return out


def op_rope(x, past, dims, mode):
    """x numpy [ne2, ne1, ne0] (C order); rotated copy returned."""
    x = _c(x).copy()
    ne2, ne1, ne0 = x.shape
    _decl("lo_op_rope", [_f32p] + [C.c_uint32] * 6)(_fp(x), ne0, ne1, ne2, past, dims, mode)
    return x


def op_scale(x, v):
    x = _c(x).copy()
    _decl("lo_op_scale", [_f32p, C.c_uint32, C.c_float])(_fp(x), x.size, v)
    return x


def op_diag_mask_inf(x, past):
    x = _c(x).copy()
    ne2, ne1, ne0 = x.shape
    _decl("lo_op_diag_mask_inf", [_f32p] + [C.c_uint32] * 4)(_fp(x), ne0, ne1, ne2, past)
    return x


def op_soft_max(x):
    x = _c(x).copy()
    x2 = x.reshape(-1, x.shape[-1])
    _decl("lo_op_soft_max", [_f32p, C.c_uint32, C.c_uint32])(_fp(x2), x2.shape[1], x2.shape[0])
    return x


def op_silu(x):
    x = _c(x)
    y = np.empty_like(x)
    _decl("lo_op_silu", [_f32p, C.c_uint32, _f32p])(_fp(x), x.size, _fp(y))
    return y


def vdot(a, b, avx: bool):
    a, b = _c(a), _c(b)
    return float(lib().lo_vdot(_fp(a), _fp(b), a.size, 1 if avx else 0))


# ----------------------------------------------------------------------------- greedy driver
def greedy_stream(octx: OracleContext, prompt_ids, predict: int, ctx_size: int, penalty: float = 1.10,
                  return_logits: bool = False):
    """Token stream the reference's generate loop produces at --temp 1e-6
    (pkg/server/server.go:110-237 + pkg/llama/llama.go:455-707): one Eval on the whole prompt,
    then per token: repetition-penalise every id present in the last-`ctx` ring (pre-filled
    with id 0, server.go:135-138) — logit<0 ? logit*1.1 : logit/1.1 (llama.go:515-522) — argmax,
    Eval on that single token.  The predict-th sampled token is never evaluated."""
    ring = [0] * ctx_size
    for t in prompt_ids:
        ring = ring[1:] + [int(t)]
    past = 0
    logits = octx.eval(prompt_ids, 0)
    past += len(prompt_ids)
    out, all_logits, margins = [], [], []
    for step in range(predict):
        all_logits.append(logits.copy())
        pen = logits.astype(np.float32) * np.float32(1e6)  # scale = float32(1/temp), llama.go:500
        present = np.zeros(pen.size, bool)
        present[np.unique(np.asarray(ring, np.int64))] = True
        neg = pen < 0
        pen = np.where(present & neg, pen * np.float32(penalty), pen)
        pen = np.where(present & ~neg, pen / np.float32(penalty), pen)
        order = np.argsort(-pen, kind="stable")
        tok = int(order[0])
        margins.append(float(pen[order[0]] - pen[order[1]]) * 1e-6)
        out.append(tok)
        ring = ring[1:] + [tok]
        if step + 1 < predict:
            logits = octx.eval([tok], past)
            past += 1
    if return_logits:
        return out, np.stack(all_logits), margins
    return out


# ----------------------------------------------------------------------------- sampler + server.Do loop
def _splitmix64(x: int) -> int:
    m = (1 << 64) - 1
    z = (x + 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


def sample_candidates(logits, last_n_tokens, top_k: int = 40, top_p: float = 0.95, temp: float = 0.8, penalty: float = 1.10):
    """SampleTopPTopK up to (not including) the random pick, pkg/llama/llama.go:498-634, in the reference's own
    arithmetic: FP32 scale / penalty (:500-527), descending sort (ties: lower id first — the reference's sort.Slice is
    unstable, any tie order is legal), top-k cut (:565), p = f32(exp(f64(v - max))), f64 sequential sum, p /= f32(sum)
    (:579-599), sequential FP32 cumsum cut at top_p and renormalisation by f32(1/cumsum) (:614-629).
    Returns (ids uint32[n], probs float32[n])."""
    lg = np.asarray(logits, np.float32)
    scale = np.float32(1.0) / np.float32(temp)
    pen = lg * scale
    present = np.zeros(lg.size, bool)
    ln = np.asarray(last_n_tokens, np.int64)
    present[ln[ln < lg.size]] = True
    neg = lg < 0
    pen = np.where(present & neg, pen * np.float32(penalty), pen).astype(np.float32)
    pen = np.where(present & ~neg, (lg * scale) / np.float32(penalty), pen).astype(np.float32)
    order = np.argsort(-pen, kind="stable")[:top_k]
    vals = pen[order]
    maxl = vals[0]
    probs = np.empty(len(order), np.float32)
    s = 0.0
    for i, v in enumerate(vals):
        p = float(np.exp(np.float64(np.float32(v - maxl))))
        probs[i] = np.float32(p)
        s += p
    probs = (probs / np.float32(s)).astype(np.float32)
    n = len(order)
    if top_p < 1.0:
        cumsum = np.float32(0.0)
        for i in range(n):
            cumsum = np.float32(cumsum + probs[i])
            if cumsum >= np.float32(top_p):
                n = i + 1
                break
        inv = np.float32(np.float32(1.0) / cumsum)
        probs = (probs[:n] * inv).astype(np.float32)
    return order[:n].astype(np.uint32), probs


def sample_pick(ids, probs, seed: int) -> int:
    """The pick of llama.go:655-673 — argmax_i p_i*p_i*f_i*f_i, f_i = float32(Int63)/2^63 — with Int63 drawn from
    splitmix64(seed + i) instead of the reference's time-seeded generator (which nothing can reproduce)."""
    best, idx = None, 0
    for i, p in enumerate(probs):
        f = np.float32(np.float32(_splitmix64((seed + i) & ((1 << 64) - 1)) >> 1) * np.float32(2.0 ** -63))
        v = np.float32(np.float32(np.float32(p * p) * f) * f)
        if best is None or v > best:
            best, idx = v, i
    return int(ids[idx])


def context_swap(ctx_size: int, keep: int, history, past: int, embd):
    """server.go:165-172.  history = last-N ids oldest first."""
    embd = list(embd)
    if past + len(embd) > ctx_size:
        left = past - keep
        past = keep
        n = left // 2
        embd = [int(t) for t in history[len(history) - n:]] + embd if n else embd
    return past, embd


def generate_stream(octx: OracleContext, prompt_ids, predict: int, ctx_size: int, top_k: int = 40, top_p: float = 0.95,
                    temp: float = 1e-6, penalty: float = 1.10, keep: int = 0, batch: int | None = None, seed: int = 0):
    """The generate loop of pkg/server.Do (server.go:127-237) on the oracle: prompt in batches, context swap, one
    SampleTopPTopK per generated token.  Returns the sampled ids."""
    batch = batch or ctx_size
    ring = [0] * ctx_size            # oldest first; append = drop the oldest (container/ring of size CtxSize, zero-filled)
    embd, out = [], []
    past = consumed = 0
    logits = None
    while len(out) < predict:
        if embd:
            past, embd = context_swap(ctx_size, keep, ring, past, embd)
            logits = octx.eval(embd, past)
        past += len(embd)
        embd = []
        if consumed < len(prompt_ids):
            while consumed < len(prompt_ids) and len(embd) < batch:
                embd.append(int(prompt_ids[consumed]))
                ring = ring[1:] + [int(prompt_ids[consumed])]
                consumed += 1
        else:
            ids, probs = sample_candidates(logits, ring, top_k, top_p, temp, penalty)
            tok = sample_pick(ids, probs, seed + len(out))
            ring = ring[1:] + [tok]
            embd.append(tok)
            out.append(tok)
    return out
