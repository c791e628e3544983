"""Model check of the fused stage hand-off (kernels_ring.cu / kernels_mega.cu, pipeline.cpp): G pipeline
stages as threads, per (stage, sequence) one buffer `x` and the flag words {in_flag, ack, seq}.  A stage step

    seq = flags.seq
    if not first:  wait flags.in_flag >= seq + 1          (the upstream stage has stored step `seq`'s residual into my x)
    if not last:   wait flags.ack     >= seq              (the downstream stage has consumed what I sent for step seq - 1)
    y = layers(x)            (the last stage overwrites its own x with intermediate layer outputs, like the kernel does)
    if not last:   downstream.x = y ; downstream.in_flag = seq + 1
    if not first:  upstream.ack = seq + 1
    flags.seq = seq + 1

must deliver every step's value through all stages exactly once, in order, for any interleaving — and the two things that
broke it on real GPUs (profiles/README.md r02r-t) must be excluded by construction: a launch outside the protocol (the graph
capture's eager warm-up) may only run before any peer decodes, and a prefill may not overlap a peer's decode."""
import random
import threading
import time

import pytest


class Ctx:
    def __init__(self):
        self.x = None
        self.in_flag = 0
        self.ack = 0
        self.seq = 0
        self.cv = threading.Condition()


def run_pipeline(G, S, steps, seed, warmup_inside_decode=False):
    rng = random.Random(seed)
    ctx = [[Ctx() for _ in range(S)] for _ in range(G)]
    out = [[None] * steps for _ in range(S)]
    errors = []
    start = threading.Barrier(G)

    def wait(c, pred):
        with c.cv:
            if not c.cv.wait_for(pred, timeout=5.0):
                raise TimeoutError("deadlock")

    def notify(c, fn):
        with c.cv:
            fn()
            c.cv.notify_all()

    def stage(g):
        try:
            first, last = g == 0, g == G - 1
            delays = random.Random(seed * 131 + g)
            if not warmup_inside_decode:
                start.wait()                      # enable_p2p's barrier: graphs (and their warm-up launches) are done everywhere
            for k in range(steps):
                for s in range(S):
                    c = ctx[g][s]
                    if warmup_inside_decode and k == 0 and s == 0 and not first:
                        wait(c, lambda: c.in_flag >= 1)   # a slow stage capturing its graph inside the first decode call, after its upstream has deposited step 0 ...
                        c.x = ("garbage", g)              # ... whose eager warm-up overwrites x outside the protocol
                    seq = c.seq
                    assert seq == k
                    if not first:
                        wait(c, lambda: c.in_flag >= seq + 1)
                    if not last:
                        wait(c, lambda: c.ack >= seq)
                    x = (s, k, 0) if first else c.x
                    if delays.random() < 0.3:
                        time.sleep(delays.random() * 0.002)
                    y = (x[0], x[1], x[2] + 1) if x[0] != "garbage" else x
                    if last:
                        c.x = ("scratch", g)      # the last stage's layers overwrite its own x
                        out[s][k] = y
                    else:
                        d = ctx[g + 1][s]
                        d.x = y
                        notify(d, lambda: setattr(d, "in_flag", seq + 1))
                    if not first:
                        u = ctx[g - 1][s]
                        notify(u, lambda: setattr(u, "ack", seq + 1))
                    c.seq = seq + 1
        except Exception as e:  # noqa: BLE001
            errors.append((g, repr(e)))

    th = [threading.Thread(target=stage, args=(g,)) for g in range(G)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=30)
    return out, errors, rng


@pytest.mark.parametrize("G,S", [(2, 1), (2, 2), (4, 4), (8, 8)])
def test_every_step_arrives_once_in_order(G, S):
    steps = 12
    out, errors, _ = run_pipeline(G, S, steps, seed=G * 100 + S)
    assert not errors, errors
    for s in range(S):
        assert out[s] == [(s, k, G) for k in range(steps)]


def test_warmup_inside_decode_is_the_bug_the_import_time_capture_removes():
    """The r02s failure, reproduced in the model: a stage that runs a launch outside the protocol after its upstream stage
    has started decoding loses the step-0 residual."""
    out, errors, _ = run_pipeline(2, 1, 4, seed=7, warmup_inside_decode=True)
    assert not errors, errors
    assert out[0][0] == ("garbage", 1) and out[0][1:] == [(0, k, 2) for k in range(1, 4)]
