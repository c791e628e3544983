"""llama.go_b200 — B200-native LLaMA forward-pass engine behind the API of gotzmann/llama.go's
pkg/ml (Tensor / op constructors / Graph / GraphCompute) and pkg/llama (NewContext / Eval).

Layout:
  csrc/      hand-written sm_100a CUDA kernels, the C++ host mirror of pkg/ml + pkg/llama and the
             C-ABI (include/llamab200.h) -> libllamab200.so
  _capi.py   ctypes binding of the C-ABI (fails loudly without the .so or without a GPU)
  ml.py      pkg/ml mirror          llama.py   pkg/llama mirror
  synth.py   synthetic weights (counter-based RNG shared with the device) + ggjt v1 reader/writer
Import as `llama_go_b200` (shim at the repo root; the directory name is not an identifier).
"""
from . import synth  # noqa: F401

__all__ = ["synth", "ml", "llama", "_capi"]
