"""Multi-GPU layer sharding.  CPU: host-side logic under a world_size-2 gloo group.
GPU: stage chaining on one device, and (when >= 2 GPUs are visible) the real NCCL pipeline."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_case


# ------------------------------------------------------------------------------------ CPU (gloo)
def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import pipeline
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the id broadcast plumbing, with a stand-in id (ncclGetUniqueId itself needs the GPU box)
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(buf, src=0)
    parts = pipeline.partition_layers(32, world)
    t = torch.tensor([float(rank + 1) * 1.5], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    q.put((rank, bytes(buf.tolist()), parts[rank], float(t.item())))
    dist.destroy_process_group()


def test_control_plane_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == bytes(range(128))          # every rank got rank 0's id
    assert res[0][2] == (0, 16) and res[1][2] == (16, 32)        # contiguous layer ranges
    assert res[0][3] == res[1][3] == 3.0                         # max over ranks


def test_partition_layers():
    sys.path.insert(0, ROOT)
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import pipeline
    for L, G in [(32, 1), (32, 2), (32, 4), (32, 8), (40, 4), (80, 8), (60, 8), (3, 2)]:
        parts = pipeline.partition_layers(L, G)
        assert parts[0][0] == 0 and parts[-1][1] == L and len(parts) == G
        assert all(parts[i][1] == parts[i + 1][0] for i in range(G - 1))
        sizes = [e - b for b, e in parts]
        assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    with pytest.raises(ValueError):
        pipeline.partition_layers(2, 3)


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_stage_chaining_on_one_gpu_equals_single_model(synth):
    """Two stage models (layers [0,1) and [1,3)) on ONE device, residual handed over by device
    pointer through lb_eval_stage: must reproduce the single-model golden logits."""
    import ctypes as C
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import _capi, llama
    lib = _capi.lib()
    rec, g = load_case("hd128")
    hp = synth.HParams(*rec["hparams"])
    tensors = list(synth.synth_model(rec["seed"], hp))
    m0 = llama.Model(hp, 0, 0, 1).load(tensors)      # tensors of other stages are accepted and ignored
    m1 = llama.Model(hp, 0, 1, 3).load(tensors)
    c0, c1 = llama.NewContext(m0, rec["context"]), llama.NewContext(m1, rec["context"])
    ids = np.ascontiguousarray(g["prompt_ids"], np.uint32)
    u32p, f32p = C.POINTER(C.c_uint32), C.POINTER(C.c_float)
    h0 = lib.lb_context_hidden_buffer(c0._h)
    logits = np.empty(hp.vocab, np.float32)

    def step(tokens, past):
        t = np.ascontiguousarray(tokens, np.uint32)
        _capi.check(lib.lb_eval_stage(c0._h, t.ctypes.data_as(u32p), t.size, past, None, None, None))
        _capi.check(lib.lb_eval_stage(c1._h, None, t.size, past, h0, None, logits.ctypes.data_as(f32p)))
        return logits.copy()

    ref = g["prompt_all_logits"][-1]
    got = step(ids, 0)
    assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max()
    past = len(ids)
    for i, tok in enumerate(g["gen_ids"][:5]):
        got = step([int(tok)], past)
        past += 1
        ref = g["step_logits"][i + 1]
        assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max()
    with pytest.raises(_capi.LlamaB200Error):
        _capi.check(lib.lb_eval_stage(c1._h, None, 1, past, None, None, None))   # stage > 0 needs hidden_in


@pytest.mark.gpu
def test_pipeline_api_single_stage_equals_eval(synth):
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import pipeline
    rec, g = load_case("tiny")
    hp = synth.HParams(*rec["hparams"])
    st = pipeline.Stage(hp, 0, 1, 0, rec["context"], 3, seed=None, tensors=synth.synth_model(rec["seed"], hp))
    ids = g["prompt_ids"]
    st.prefill(np.stack([ids] * 3), 0)
    gen = np.asarray(g["gen_ids"][:-1], np.uint32)
    st.decode(np.stack([gen] * 3), len(ids))
    ref = g["step_logits"][len(gen)]
    for s in range(3):
        assert np.abs(st.logits(s) - ref).max() <= 1e-3 * np.abs(ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("handoff", ["nccl", "p2p"])
def test_pipeline_two_gpus(handoff):
    """Two stages on two GPUs, one process each: golden logits through the NCCL send/recv hand-off and through the
    hand-off fused into the stage kernels (NVLink peer stores + flags)."""
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import _capi
    if _capi.lib().lb_device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if handoff == "nccl" else "29534", os.path.join(ROOT, "tests", "mgpu_worker.py"), "hd128", handoff]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, env=env)
    out = p.stdout.decode("utf-8", "replace")
    assert p.returncode == 0 and "MGPU_OK" in out, out[-3000:]
    assert ("hand-off: p2p-fused" in out) == (handoff == "p2p"), out[-3000:]
