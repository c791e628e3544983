import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["tiny", "hd128", "wide3h", "long"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_case(name):
    """(record dict, npz) of a committed golden case (tests/golden, made by tools/gen_golden.py)."""
    with open(os.path.join(GOLDEN, f"refbin_{name}.json")) as f:
        rec = json.load(f)
    npz = np.load(os.path.join(GOLDEN, f"logits_{name}.npz"))
    return rec, npz


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (test infrastructure; see oracle/llama_oracle.c)."""
    from oracle import oracle as O
    O.build()
    O.set_dot_mode(False)
    return O


@pytest.fixture(scope="session")
def synth():
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import synth as S
    return S
