#!/usr/bin/env python
"""Per-phase timing of the pod-batch ring megakernel (CTA 0's globaltimer stamps), LLaMA-7B FP32, B pods.
   python tools/pods_trace.py [--pods 8] [--past 400]"""
import argparse, ctypes as C, os, sys
import numpy as np
os.environ["LB_MEGA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_go_b200  # noqa
from llama_go_b200 import _capi, llama, synth

ap = argparse.ArgumentParser(); ap.add_argument("--pods", type=int, default=8); ap.add_argument("--past", type=int, default=400)
ap.add_argument("--layers", type=int, default=32)
a = ap.parse_args()
hp = synth.HParams(32000, 4096, 256, 32, a.layers)
m = llama.Model(hp).init_random(0)
rs = np.random.RandomState(0)
pods = [llama.NewContext(m, 512) for _ in range(a.pods)]
for c in pods:
    llama.Eval(c, rs.randint(3, hp.vocab, size=a.past).astype(np.uint32), 0)
b = llama.PodBatch(pods)
for i in range(4):
    b.Eval(rs.randint(3, hp.vocab, size=a.pods).astype(np.uint32), [a.past + i] * a.pods)
n = a.layers * 13
buf = (C.c_uint64 * n)()
_capi.check(_capi.lib().lb_batch_mega_trace(b._h, buf, n))
t = np.array(buf[:], dtype=np.int64).reshape(a.layers, 13)
d = np.diff(t, axis=1)[1:-1]
names = ["rms scales", "gemv qkv", "barrier1", "attention", "barrier2", "gemv wo", "barrier3", "rms scales 2", "gemv w1w3", "barrier4", "gemv w2", "barrier5"]
ideal = {"gemv qkv": 201.4e6, "gemv wo": 67.2e6, "gemv w1w3": 360.8e6, "gemv w2": 180.4e6}
print(f"pods = {a.pods}; phase mean_us min_us max_us (CTA 0 view; ideal at 7.0 TB/s)")
for i, nme in enumerate(names):
    extra = f"   ideal {ideal[nme] / 7.0e12 * 1e6:6.1f}" if nme in ideal else ""
    print(f"{nme:14s} {d[:, i].mean() / 1e3:8.2f} {d[:, i].min() / 1e3:8.2f} {d[:, i].max() / 1e3:8.2f}{extra}")
print(f"layer total    {np.diff(t[:, [0, 12]], axis=1)[1:-1].mean() / 1e3:8.2f} us")
