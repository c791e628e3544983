#!/usr/bin/env python
"""Why is a host-synchronous step of the ring megakernel slower than a back-to-back one?  Times, on 7B FP32:
   (a) DecodeResident of 20 steps (one graph replay after the other), (b) 20 x DecodeResident of 1 step with a host
   sync in between, (c) 20 x lb_eval (host buffers).   LB_RING=1 python tools/ring_sync_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_go_b200  # noqa
from llama_go_b200 import _capi, llama, synth
hp = synth.HParams(32000, 4096, 256, 32, 32)
m = llama.Model(hp).init_random(0)
c = llama.NewContext(m, 512)
rs = np.random.RandomState(0)
llama.Eval(c, rs.randint(3, hp.vocab, size=384).astype(np.uint32), 0)
gen = rs.randint(3, hp.vocab, size=64).astype(np.uint32)
lib = _capi.lib()
for rep in range(3):
    llama.DecodeResident(c, gen[:20], 384)
    a = llama.DecodeResident(c, gen[:20], 384) / 20
    lib.lb_context_synchronize(c._h)
    b_dev = []
    t0 = time.perf_counter()
    for i in range(20):
        b_dev.append(llama.DecodeResident(c, gen[i:i + 1], 384 + i))
    b_wall = (time.perf_counter() - t0) / 20 * 1e3
    t0 = time.perf_counter()
    for i in range(20):
        llama.Eval(c, gen[i:i + 1], 384 + i)
    c_wall = (time.perf_counter() - t0) / 20 * 1e3
    t0 = time.perf_counter()
    for i in range(20):
        llama.Eval(c, gen[i:i + 1], 384 + i)
        time.sleep(0.002)
    d_wall = (time.perf_counter() - t0) / 20 * 1e3 - 2.0
    print(f"rep {rep}: back-to-back {a:.3f} ms/step | 1-step resident: device {np.mean(b_dev):.3f} (min {min(b_dev):.3f} max {max(b_dev):.3f}) wall {b_wall:.3f} | lb_eval wall {c_wall:.3f} | lb_eval + 2 ms idle: {d_wall:.3f}", flush=True)
