"""Multi-GPU layer sharding host logic: one process per GPU, rank r owns a
contiguous layer range; the residual stream moves between ranks by NCCL send/recv inside the
C-ABI (lb_pipeline_prefill / lb_pipeline_decode).  torch.distributed is used for the control
plane only (NCCL unique-id broadcast, barriers, max-over-ranks timing)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi, llama
from ._capi import check, lib

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)


def partition_layers(layers: int, world: int):
    """[(begin, end)] per rank: an even contiguous split (remainder to the FIRST ranks).  The lm_head
    (0.65 layer-equivalents for 7B) rides on the last rank; for L in {32,40,60,80} and G in {2,4,8}
    the even split minimises the slowest stage."""
    if not (1 <= world <= layers):
        raise ValueError("need 1 <= world <= layers")
    base, rem = divmod(layers, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < rem else 0)
        out.append((b, e))
        b = e
    return out


def exchange_unique_id(rank: int, dist=None) -> bytes:
    """rank 0 creates the NCCL unique id, everybody receives it (torch.distributed broadcast)."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = (C.c_ubyte * 128)()
        check(lib().lb_comm_unique_id(raw))
        buf = torch.tensor(list(raw), dtype=torch.uint8)
    if dist is not None:
        dist.broadcast(buf, src=0)
    return bytes(buf.tolist())


class Stage:
    """This rank's slice of the model plus one llama.Context per in-flight sequence."""

    def __init__(self, hp, rank: int, world: int, device: int, ctx_size: int, n_seq: int, seed: int | None = 0,
                 tensors=None):
        self.hp, self.rank, self.world = hp, rank, world
        self.begin, self.end = partition_layers(hp.layers, world)[rank]
        self.model = llama.Model(hp, device, self.begin, self.end)
        if tensors is not None:
            self.model.load(tensors)
        elif seed is not None:
            self.model.init_random(seed)
        self.ctxs = [llama.NewContext(self.model, ctx_size) for _ in range(n_seq)]
        self._arr = (C.c_void_p * n_seq)(*[c._h for c in self.ctxs])
        self.p2p = False
        self._dist = None

    @property
    def is_first(self):
        return self.begin == 0

    @property
    def is_last(self):
        return self.end == self.hp.layers

    def prefill(self, tokens, past: int = 0):
        """tokens [n_seq][n] (read on rank 0 only; other ranks may pass shape-compatible zeros)."""
        t = np.ascontiguousarray(tokens, dtype=np.uint32)
        check(lib().lb_pipeline_prefill(self._arr, len(self.ctxs), t.ctypes.data_as(_u32p), t.shape[1], past))
        if self.p2p and self._dist is not None:
            # the prefill is outside the decode hand-off's flag protocol: an upstream stage that is already decoding would store
            # step 0's residual into an x buffer this stage's prefill kernels are still reading.  lb_pipeline_prefill returns
            # with this rank's work complete; nobody decodes before every rank is here.
            self._dist.barrier()

    def decode(self, tokens, past: int) -> float:
        """tokens [n_seq][steps]; returns CUDA-event ms on this rank."""
        t = np.ascontiguousarray(tokens, dtype=np.uint32)
        ms = C.c_float(0)
        check(lib().lb_pipeline_decode(self._arr, len(self.ctxs), t.ctypes.data_as(_u32p), t.shape[1], past, C.byref(ms)))
        return ms.value

    def enable_p2p(self, dist) -> bool:
        """Fuse the stage hand-off into the stage kernels (NVLink peer stores + flags, lb_pipeline_p2p_*): exchange the CUDA
        IPC handles of every rank's contexts and map the neighbours'.  All ranks switch together or not at all (any failure ->
        NCCL send/recv as before).  LB_PIPE_NCCL=1 keeps NCCL.  Call before the first decode()."""
        import os
        import torch
        if self.world == 1 or os.environ.get("LB_PIPE_NCCL"):
            return False
        n = len(self.ctxs)
        ok, mine = 1.0, b""
        try:
            buf = (C.c_ubyte * (128 * n))()
            check(lib().lb_pipeline_p2p_export(self._arr, n, buf))
            mine = bytes(buf)
        except Exception:
            ok = 0.0
        allh = [None] * self.world
        dist.all_gather_object(allh, mine)
        if ok and all(len(h) == 128 * n for h in allh):
            try:
                down = allh[self.rank + 1] if self.rank + 1 < self.world else None
                up = allh[self.rank - 1] if self.rank > 0 else None
                check(lib().lb_pipeline_p2p_import(self._arr, n, down, up))
            except Exception:
                ok = 0.0
        else:
            ok = 0.0
        t = torch.tensor([ok])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        self.p2p = bool(t.item() == 1.0)
        self._dist = dist if self.p2p else None
        if not self.p2p:
            try:
                lib().lb_pipeline_p2p_disable(self._arr, n)
            except Exception:
                pass
        return self.p2p

    def logits(self, seq: int) -> np.ndarray:
        return llama.ReadLogits(self.ctxs[seq]).copy()

    def free(self):
        """Release the contexts, then the stage model (device memory) now rather than at garbage collection."""
        for c in self.ctxs:
            c.ReleaseContext()
        self.ctxs = []
        self.model.free()
