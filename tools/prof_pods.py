#!/usr/bin/env python
"""A short pod-batch run for ncu: 7B-shaped model with few layers, B pods, a few steps.
   python tools/prof_pods.py [--layers 4] [--pods 8] [--steps 3]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_go_b200  # noqa
from llama_go_b200 import llama, synth
ap = argparse.ArgumentParser(); ap.add_argument("--layers", type=int, default=4); ap.add_argument("--pods", type=int, default=8)
ap.add_argument("--steps", type=int, default=3); ap.add_argument("--prompt", type=int, default=384)
a = ap.parse_args()
hp = synth.HParams(32000, 4096, 256, 32, a.layers)
m = llama.Model(hp).init_random(0)
rs = np.random.RandomState(0)
pods = [llama.NewContext(m, 512) for _ in range(a.pods)]
for c in pods:
    llama.Eval(c, rs.randint(3, hp.vocab, size=a.prompt).astype(np.uint32), 0)
b = llama.PodBatch(pods)
gen = rs.randint(3, hp.vocab, size=(a.pods, a.steps)).astype(np.uint32)
ms = b.DecodeResident(gen, [a.prompt] * a.pods)
print("ms per step", ms / a.steps)
