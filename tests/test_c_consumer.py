"""A plain-C consumer of include/llamab200.h (tests/c_consumer/consumer.c), compiled with gcc -std=c11 and
linked against libllamab200.so the way cgo would bind it — the Go shim of INTEGRATION.md cannot be compiled
here (no Go toolchain), so this proves the header is valid C and the ABI links and runs from C.
Reference seam: pkg/llama/llama.go:91-113 (NewContext), 211-218 (Eval)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_consumer", "consumer.c")
LIBDIR = os.path.join(ROOT, "llama.go_b200")


def build_consumer(tmpdir):
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(LIBDIR, "libllamab200.so")):
        g.build()
    exe = os.path.join(str(tmpdir), "consumer")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           SRC, "-o", exe, "-L", LIBDIR, "-lllamab200", "-Wl,-rpath," + LIBDIR])
    return exe


def test_header_is_valid_c11_and_links(tmp_path):
    exe = build_consumer(tmp_path)
    # the header alone, as strict C (what cgo's preamble compile does)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c",
                           os.path.join(ROOT, "include", "llamab200.h")])
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert out.returncode == 0, out.stderr.decode()
    assert out.stdout.decode().startswith("version ")


@pytest.mark.gpu
def test_c_consumer_runs_eval_and_matches_the_ctypes_path(tmp_path):
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import llama, synth
    exe = build_consumer(tmp_path)
    out = subprocess.run([exe, "7"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert out.returncode == 0, out.stderr.decode()
    lines = out.stdout.decode().split("\n")
    assert "OK" in lines
    got_p = np.array([float.fromhex(x[2:]) for x in lines if x.startswith("P ")], np.float32)
    got_d = np.array([float.fromhex(x[2:]) for x in lines if x.startswith("D ")], np.float32)
    hp = synth.HParams(96, 64, 32, 2, 2)
    model = llama.Model(hp).init_random(7)
    model.set_tensor("norm.weight", (1.0 + np.arange(64) / 64.0).astype(np.float32))
    lctx = llama.NewContext(model, 32)
    ref_p = llama.Eval(lctx, [1, 35, 36, 90, 7], 0).copy()
    ref_d = llama.Eval(lctx, [11], 5).copy()
    np.testing.assert_array_equal(got_p, ref_p)     # same library, same kernels: bit-identical
    np.testing.assert_array_equal(got_d, ref_d)
