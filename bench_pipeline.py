"""Multi-GPU arm of bench.py: LLaMA-7B FP32 decode, layer-sharded over N GPUs (one process per GPU,
launched by torchrun), N sequences in flight, NCCL send/recv of the residual between stages.

A step = every in-flight sequence advances one token (N tokens per step, "scaling": "weak": each GPU
streams 1/N of the weights N times per step = the same bytes per step as the single-GPU arm)."""
import ctypes as C
import json
import os
import time

import numpy as np

from bench import CTX, MODELS, PROMPT_LEN, UNIT, ClockSampler, measured_peak, metric_name, rank_world


def pipeline_parity(stage, hp, device, ctx_size, prompt, gen, n_decoded, S):
    """Single-GPU reference of the pipelined run on the last rank's GPU: unsharded model (same device RNG, seed 0),
    prompt prefill + `n_decoded` teacher-forced single-token steps per sequence; returns the worst relative error
    of the last step's logits over the sequences.  None when the unsharded model does not fit."""
    from llama_go_b200 import llama
    full_bytes = 4 * (hp.layers * (4 * hp.dim ** 2 + 3 * hp.dim * hp.ff + 2 * hp.dim) + 2 * hp.vocab * hp.dim + hp.dim)
    stage_bytes = full_bytes * (stage.end - stage.begin) / hp.layers + 4 * hp.vocab * hp.dim
    kv_bytes = 2 * 4 * hp.layers * ctx_size * hp.dim
    if full_bytes + kv_bytes + stage_bytes + S * kv_bytes * (stage.end - stage.begin) / hp.layers + 12e9 > 178e9:
        return {"rel_err": None}
    full = llama.Model(hp, device).init_random(0)
    worst = 0.0
    c = llama.NewContext(full, ctx_size)
    for s in range(S):
        llama.Eval(c, prompt[s], 0)
        llama.DecodeResident(c, gen[s, :n_decoded], prompt.shape[1])
        ref = llama.ReadLogits(c).astype(np.float64)
        got = stage.logits(s).astype(np.float64)
        worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
    c.ReleaseContext()
    return {"rel_err": worst}


def pipeline_measure(args, hp, model_key, ctx_size, K, W, env):
    """One layer-sharded measurement on the already initialised process group / NCCL communicator.  Returns the record
    (meaningful on rank 0) — value (device timed, max over ranks), e2e, parity vs the unsharded model, roofline."""
    torch, dist, lib, pipeline, rank, world, local = env
    from llama_go_b200 import _capi
    S = world                                   # sequences in flight
    t_setup = time.time()
    stage = pipeline.Stage(hp, rank, world, local, ctx_size, S, seed=0)
    rs = np.random.RandomState(0)
    prompt = rs.randint(3, hp.vocab, size=(S, PROMPT_LEN)).astype(np.uint32)
    gen = rs.randint(3, hp.vocab, size=(S, 2 * W + 2 * K)).astype(np.uint32)
    p2p = stage.enable_p2p(dist)                # fused NVLink hand-off in the stage kernels (else NCCL send/recv)
    stage.prefill(prompt, 0)                    # setup, untimed
    t_setup = time.time() - t_setup

    def sync_all():
        for c in stage.ctxs[:1]:
            _capi.check(lib.lb_context_synchronize(c._h))
        dist.barrier()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: K pipelined steps, device timed, max over ranks
    stage.decode(gen[:, :W], PROMPT_LEN)      # warm-up: graphs captured, NCCL p2p channels connected
    sync_all()
    stage.decode(gen[:, :W], PROMPT_LEN)      # second warm-up pass (one run was seen 2x slow on a cold pair of GPUs)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.lb_kernel_launches()
    ms = stage.decode(gen[:, W:W + K], PROMPT_LEN + W)
    sync_all()
    launches = lib.lb_kernel_launches() - l0
    ms = max_over_ranks(ms)
    value = S * K / (ms / 1e3)

    # ---- e2e: host-driven, one synchronous pipelined step at a time (token ids H2D on rank 0,
    #      logits of every sequence D2H on the last rank, inside the timed region)
    past = PROMPT_LEN + W + K
    for i in range(W):
        stage.decode(gen[:, W + K + i:W + K + i + 1], past + i)
    sync_all()
    t0 = time.perf_counter()
    for i in range(K):
        stage.decode(gen[:, 2 * W + K + i:2 * W + K + i + 1], past + W + i)
        if stage.is_last:
            for s in range(S):
                stage.logits(s)
    sync_all()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    e2e = S * K / e2e_s

    # ---- parity (untimed): the last stage's logits of every in-flight sequence after the LAST pipelined step must
    #      equal a single-GPU evaluation of the same token history (prompt + every teacher-forced token of the
    #      four decode passes above) by the unsharded model with the same synthetic weights.
    parity = pipeline_parity(stage, hp, local, ctx_size, prompt, gen, 2 * W + 2 * K, S) if stage.is_last else None
    par_t = torch.tensor([-1.0 if parity is None or parity["rel_err"] is None else parity["rel_err"]], dtype=torch.float64)
    dist.all_reduce(par_t, op=dist.ReduceOp.MAX)
    dist.barrier()

    all_launches = torch.tensor([float(launches)], dtype=torch.float64)
    dist.all_reduce(all_launches, op=dist.ReduceOp.SUM)
    stage.free()
    dist.barrier()
    peak, peak_src = measured_peak()
    T_mid = PROMPT_LEN + W + K / 2.0
    wbytes = 4 * (hp.layers * (4 * hp.dim ** 2 + 3 * hp.dim * hp.ff + 2 * hp.dim) + hp.vocab * hp.dim + 2 * hp.dim)
    bytes_per_token = wbytes + 2 * hp.layers * T_mid * hp.dim * 4 + 2 * hp.layers * hp.dim * 4 + 4 * hp.vocab
    agg_gbs = bytes_per_token * value / 1e9
    have_par = par_t.item() >= 0
    return {
        "metric": metric_name(model_key), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LLaMA-%s FP32 decode, context %d, %d-token prompts prefilled, %d sequences in flight"
                               % (model_key.upper(), ctx_size, PROMPT_LEN, S),
                   "parallelism": "pp%d (layer-sharded, %d layers/GPU, %s)" % (
                       world, hp.layers // world,
                       "residual handed to the next stage by the stage kernel itself: NVLink peer stores + flag (CUDA IPC), NCCL for the prefill"
                       if p2p else "NCCL send/recv of the residual"),
                   "handoff": "p2p-fused" if p2p else "nccl",
                   "sequences_in_flight": S, "tokens_per_step": S, "weights": "random-init (device RNG, seed 0)",
                   "kv_cache": "fp32 in HBM", "l2": "inputs>L2", "setup_s": round(t_setup, 1),
                   "note": "a single sequence gains nothing from layer sharding (dependency chain); "
                           "throughput is aggregate over the in-flight sequences (the reference's pods)"},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 4 * S + 8 * S, "d2h_bytes_per_step": 4 * hp.vocab * S,
                "api": "lb_pipeline_decode(steps=1) + lb_context_read_logits per sequence, synchronous per step"},
        "gpu_launches": int(all_launches.item()),
        "roofline": {"bound": "hbm", "kernel": "whole step (aggregate over GPUs)", "achieved": round(agg_gbs, 1),
                     "peak": peak * world, "unit": "GB/s", "frac": round(agg_gbs / (peak * world), 4), "traffic": None,
                     "peak_source": peak_src + " x n_gpus"},
        "cpu_baseline": None,
        "parity_rel_err": (float(par_t.item()) if have_par else None),
        "parity": ("max over the in-flight sequences of max|logits_pipeline - logits_single_gpu| / max|logits_single_gpu| "
                   "after the last step (single-GPU lb_eval + lb_decode_resident of the same tokens, unsharded model on the "
                   "last rank's GPU)" if have_par else "skipped: the unsharded model does not fit beside this stage on one GPU"),
    }


def run_pipeline(args):
    import torch
    import torch.distributed as dist
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import _capi, pipeline, synth

    rank, world, local = rank_world()
    # stdout carries exactly one JSON line: NCCL prints "NCCL version ..." there when NCCL_DEBUG=VERSION is inherited
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _capi.require_gpu()
    lib = _capi.lib()
    uid = pipeline.exchange_unique_id(rank, dist)
    _capi.check(lib.lb_comm_init(uid, rank, world, local))
    env = (torch, dist, lib, pipeline, rank, world, local)

    hp = getattr(synth, MODELS[args.model])
    K, W = args.steps, args.warmup
    ctx_size = max(args.context or CTX, PROMPT_LEN + 2 * W + 2 * K + 2)
    line = pipeline_measure(args, hp, args.model, ctx_size, K, W, env)
    # ---- the multi-GPU BASELINE configurations next to the 7B headline: 13B on 2/4 GPUs (config 4),
    #      65B at context 2048 on 8 GPUs (config 5)
    if args.model == "7b" and not getattr(args, "no_configs", False):
        extra = {2: ("13b", "llama13b_ctx512", CTX), 4: ("13b", "llama13b_ctx512", CTX), 8: ("65b", "llama65b_ctx2048", 2048)}.get(world)
        if extra:
            key, name, cctx = extra
            try:
                rec = pipeline_measure(args, getattr(synth, MODELS[key]), key, max(cctx, PROMPT_LEN + 2 * W + 2 * K + 2), K, W, env)
            except Exception as e:   # keep the headline
                rec = {"error": str(e)}
            line["configs"] = {name: rec}
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.barrier()
    lib.lb_comm_destroy()
    dist.destroy_process_group()
