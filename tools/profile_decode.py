#!/usr/bin/env python
"""Small driver for ncu: builds a LLaMA-7B-shaped model, evaluates a short prompt, then runs a few
single-token decode steps with plain kernel launches (LB_NO_GRAPH=1) so every kernel shows up by name.

    LB_NO_GRAPH=1 ncu ... python tools/profile_decode.py [--layers 32] [--prompt 8] [--past 448] [--steps 2]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LB_NO_GRAPH", "1")
import llama_go_b200  # noqa: E402,F401
from llama_go_b200 import _capi, llama, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--prompt", type=int, default=8)
    ap.add_argument("--past", type=int, default=448, help="position the decode steps start at (KV beyond the prompt is zero)")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--ctx", type=int, default=512)
    ap.add_argument("--weights", default="f32", choices=["f32", "q8"])
    a = ap.parse_args()
    hp = synth.HParams(32000, 4096, 256, 32, a.layers)
    lib = _capi.lib()
    model = llama.Model(hp, weight_type=llama.LB_TYPE_Q8_0 if a.weights == "q8" else llama.LB_TYPE_F32).init_random(0)
    lctx = llama.NewContext(model, a.ctx)
    print("launches after init:", lib.lb_kernel_launches(), flush=True)
    rs = np.random.RandomState(0)
    llama.Eval(lctx, rs.randint(3, hp.vocab, size=a.prompt), 0)
    print("launches after prompt:", lib.lb_kernel_launches(), flush=True)
    for i in range(a.steps):
        llama.Eval(lctx, [int(rs.randint(3, hp.vocab))], a.past + i)
    print("launches after decode:", lib.lb_kernel_launches(), flush=True)


if __name__ == "__main__":
    main()
