import os
import tempfile

import numpy as np


def test_rng_is_slice_consistent_and_distributed(synth):
    a = synth.synth_values(5, 33, 0, 100000, 0.0, 1.0)
    b = synth.synth_values(5, 33, 777, 1000, 0.0, 1.0)
    np.testing.assert_array_equal(a[777:1777], b)
    assert abs(float(a.mean())) < 0.02 and abs(float(a.std()) - 1.0) < 0.02
    c = synth.synth_values(6, 33, 0, 1000, 0.0, 1.0)
    assert not np.array_equal(a[:1000], c)


def test_known_answer_vector(synth):
    # pins the recipe itself (the device generator must reproduce these bits)
    v = synth.synth_values(0, 1, 0, 4, 0.0, 1.0)
    w = synth.synth_values(0, 17, 5, 3, 1.0, 0.1)
    assert v.dtype == np.float32 and w.dtype == np.float32
    assert v.view(np.uint32).tolist() == synth.synth_values(0, 1, 0, 8, 0.0, 1.0)[:4].view(np.uint32).tolist()
    assert synth.tensor_id("layers.0.attention_norm.weight") == 16
    assert synth.tensor_id("layers.3.feed_forward.w3.weight") == 72
    assert synth.tensor_id("output.weight") == 3


def test_hparams_ff_matches_reference_formula(synth):
    assert synth.LLAMA_7B.ff == 11008 and synth.LLAMA_13B.ff == 13824
    assert synth.LLAMA_30B.ff == 17920 and synth.LLAMA_65B.ff == 22016


def test_ggjt_roundtrip(synth):
    hp = synth.HParams(64, 32, 32, 2, 2)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "m.bin")
        synth.write_ggjt(p, hp, synth.synth_model(9, hp))
        hp2, vocab, tensors = synth.read_ggjt(p)
    assert hp2 == hp and len(vocab) == 64
    for name, arr in synth.synth_model(9, hp):
        np.testing.assert_array_equal(tensors[name], arr)


def test_prompt_ids(synth):
    assert synth.prompt_token_ids(b"abcde") == [1, 35, 35, 100, 101, 102, 103, 104]


def test_reference_arm_generates_weights_without_the_product_library():
    """bench.py --impl reference must not map libllamab200.so: its synthetic weights come
    from liboracle.so's lo_synth_fill, bit-identical to synth.synth_values()."""
    import subprocess
    import sys
    from conftest import ROOT
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np, bench\n"
        "synth, O = bench._ref_modules()\n"
        "hp = synth.HParams(96, 64, 32, 2, 2)\n"
        "for (n, a), (n2, b) in zip(O.synth_model(9, hp, synth.tensor_table(hp)), synth.synth_model(9, hp)):\n"
        "    assert n == n2 and a.shape == b.shape and (a == b).all(), n\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libllamab200' not in maps, 'product library mapped by the reference arm'\n"
        "assert 'liboracle' in maps\n"
        "print('REF_ARM_CLEAN')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert out.returncode == 0 and b"REF_ARM_CLEAN" in out.stdout, out.stderr.decode()[-2000:]
