#!/usr/bin/env python
"""Per-phase timing of the decode megakernel (CTA 0's globaltimer stamps), LLaMA-7B FP32.
   LB_MEGA_TRACE=1 python tools/mega_trace.py [--past 448]"""
import argparse, ctypes as C, os, sys
import numpy as np
os.environ["LB_MEGA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_go_b200  # noqa
from llama_go_b200 import _capi, llama, synth

ap = argparse.ArgumentParser(); ap.add_argument("--past", type=int, default=448); ap.add_argument("--layers", type=int, default=32); ap.add_argument("--q8", action="store_true")
a = ap.parse_args()
hp = synth.HParams(32000, 4096, 256, 32, a.layers)
m = llama.Model(hp, weight_type=llama.LB_TYPE_Q8_0 if a.q8 else llama.LB_TYPE_F32).init_random(0)
c = llama.NewContext(m, 512)
llama.Eval(c, [5, 6, 7, 8, 9, 10, 11, 12, 13], 0)
for i in range(4):
    llama.Eval(c, [7 + i], a.past + i)
n = a.layers * 13 + 13 * 148
buf = (C.c_uint64 * n)()
_capi.check(_capi.lib().lb_context_mega_trace(c._h, buf, n))
allv = np.array(buf[:], dtype=np.int64)
t = allv[:a.layers * 13].reshape(a.layers, 13)
arr = allv[a.layers * 13:a.layers * 13 + 5 * 148].reshape(5, 148)
prod = allv[a.layers * 13 + 5 * 148:].reshape(8, 148)   # ring only: producer stall ns [4 phases], jobs [4 phases]
d = np.diff(t, axis=1)[1:-1]          # skip first/last layer
names = ["rms1", "gemv qkv", "barrier1", "attention", "barrier2", "gemv wo", "barrier3", "rms2", "gemv w1w3", "barrier4", "gemv w2", "barrier5"]
ideal = {"gemv qkv": 201.4e6, "gemv wo": 67.2e6, "gemv w1w3": 360.8e6, "gemv w2": 180.4e6}
if a.q8:
    ideal = {k: v * 36 / 128 for k, v in ideal.items()}
print("phase            mean_us   min_us   max_us   (CTA 0 view; ideal at 7.0 TB/s)")
for i, nme in enumerate(names):
    extra = f"   ideal {ideal[nme] / 7.0e12 * 1e6:6.1f}" if nme in ideal else ""
    print(f"{nme:14s} {d[:, i].mean() / 1e3:8.2f} {d[:, i].min() / 1e3:8.2f} {d[:, i].max() / 1e3:8.2f}{extra}")
print(f"layer total    {np.diff(t[:, [0, 12]], axis=1)[1:-1].mean() / 1e3:8.2f} us")

print("arrival spread of the 148 CTAs at layer 5's barriers (us after the first arrival): p50 / p90 / max")
for b, nme in enumerate(["after qkv", "after attention", "after wo", "after w1w3", "after w2"]):
    x = (arr[b] - arr[b].min()) / 1e3
    print(f"  barrier {b + 1} ({nme:15s}): {np.percentile(x, 50):6.2f} {np.percentile(x, 90):6.2f} {x.max():6.2f}   slowest CTAs: {np.argsort(-x)[:6].tolist()}")

if prod[4:].sum() > 0:
    print("TMA ring, layer 5, per CTA and MulMat phase: producer time blocked on a full ring (us) and rows taken: p10 / p50 / p90 / max")
    for i, nme in enumerate(["qkv", "wo", "w1w3", "w2"]):
        st = prod[i] / 1e3
        jb = prod[4 + i]
        print(f"  {nme:5s} stall {np.percentile(st, 10):6.2f} {np.percentile(st, 50):6.2f} {np.percentile(st, 90):6.2f} {st.max():6.2f}   rows {np.percentile(jb, 10):5.0f} {np.percentile(jb, 50):5.0f} {np.percentile(jb, 90):5.0f} {jb.max():5.0f}  (mean {jb.mean():.1f})")
    # correlation: do the CTAs that arrive last take more rows?
    for b, i in ((0, 0), (2, 1), (3, 2), (4, 3)):
        late = (arr[b] - arr[b].min()) / 1e3
        print(f"  barrier {b + 1}: corr(lateness, rows) = {np.corrcoef(late, prod[4 + i])[0, 1]:+.2f}, corr(lateness, stall) = {np.corrcoef(late, prod[i])[0, 1]:+.2f}; lateness of the 8 CTAs with most rows: {np.round(late[np.argsort(-prod[4 + i])[:8]], 2).tolist()}")
