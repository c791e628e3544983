"""The C-ABI library loads and exports every symbol include/llamab200.h declares (no GPU needed)."""
import ctypes as C
import os

import pytest


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        g.build()
    return _capi


def test_library_exports_every_declared_symbol(capi):
    lib = C.CDLL(capi.LIB_PATH)
    declared = capi.header_symbols()
    assert len(declared) >= 45
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/llamab200.h but not exported"
    # the Python binding covers the header exactly
    assert sorted(capi._SIGS) == declared


def test_no_torch_types_in_the_boundary(capi):
    import re
    with open(capi.HEADER_PATH) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)  # declarations only
    for banned in ("torch", "at::", "std::", "Tensor&"):
        assert banned not in src


def test_fails_loudly_without_a_gpu(capi):
    """There is no CPU fallback: without a usable sm_100 device every constructor refuses."""
    lib = capi.lib()
    if lib.lb_device_count() > 0:
        pytest.skip("a GPU is present")
    hp = capi.HParamsC(64, 32, 32, 2, 1)
    assert not lib.lb_model_create(C.byref(hp), 0, 0, 1, 0)
    assert b"no CUDA device" in lib.lb_last_error() or b"CPU fallback" in lib.lb_last_error()
    assert not lib.lb_ml_new_context(0)
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import llama, ml, synth
    with pytest.raises(capi.LlamaB200Error):
        llama.Model(synth.HParams(64, 32, 32, 2, 1))
    with pytest.raises(capi.LlamaB200Error):
        ml.NewContext()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under llama.go_b200/ may reference it."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama.go_b200")
    for dp, _, files in os.walk(root):
        if "build" in dp.split(os.sep):
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".go")) or fn == "Makefile":
                with open(os.path.join(dp, fn), errors="replace") as f:
                    src = f.read()
                assert "liboracle" not in src and "llama_oracle" not in src and "from oracle" not in src \
                    and "import oracle" not in src, f"{fn} references the oracle"
