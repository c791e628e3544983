"""Worker for tests/test_multi_gpu.py (launched by torchrun, one process per GPU): runs a golden
case through the layer-sharded NCCL pipeline and checks the last stage's logits against the golden
vectors (oracle pinned to the reference binary)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import _capi, pipeline, synth
    from conftest import load_case

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _capi.lib()
    uid = pipeline.exchange_unique_id(rank, dist)
    _capi.check(lib.lb_comm_init(uid, rank, world, local))
    case = sys.argv[1] if len(sys.argv) > 1 else "hd128"
    rec, g = load_case(case)
    hp = synth.HParams(*rec["hparams"])
    S = 2  # two in-flight sequences, both fed the same tokens -> both must reproduce the golden logits
    stage = pipeline.Stage(hp, rank, world, local, rec["context"], S, seed=None, tensors=synth.synth_model(rec["seed"], hp))
    ids = g["prompt_ids"]
    p2p = stage.enable_p2p(dist) if (len(sys.argv) > 2 and sys.argv[2] == "p2p") else False
    print(f"[rank {rank}] hand-off: {'p2p-fused' if p2p else 'nccl'}", flush=True)
    stage.prefill(np.stack([ids] * S), 0)
    ok = True
    if stage.is_last:
        for s in range(S):
            ref = g["prompt_all_logits"][-1]
            err = np.abs(stage.logits(s) - ref).max() / np.abs(ref).max()
            ok &= bool(err <= 1e-3)
            print(f"[rank {rank}] prefill seq {s} rel err {err:.3e}", flush=True)
    gen = np.asarray(g["gen_ids"][:-1], np.uint32)
    steps = len(gen)
    ms = stage.decode(np.stack([gen] * S), len(ids))
    if stage.is_last:
        for s in range(S):
            ref = g["step_logits"][steps]
            err = np.abs(stage.logits(s) - ref).max() / np.abs(ref).max()
            ok &= bool(err <= 1e-3)
            print(f"[rank {rank}] decode({steps} steps, {ms:.2f} ms) seq {s} rel err {err:.3e}", flush=True)
    import torch
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    lib.lb_comm_destroy()
    dist.destroy_process_group()
    if t.item() != 1.0:
        sys.exit(3)
    if rank == 0:
        print("MGPU_OK", flush=True)


if __name__ == "__main__":
    main()
