"""Native ggjt v1 loader (csrc/loader.cpp): a file written in the reference's on-disk
format loads into HBM with the same tensors (F32 bit-exact, F16 widened like llama.go:938-941) and the
same acceptance rules as pkg/llama.LoadModel."""
import os
import struct
import tempfile

import numpy as np
import pytest

from conftest import load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import llama
    return llama


@pytest.mark.parametrize("f16", [False, True])
def test_load_ggjt_matches_set_tensor_path(L, synth, f16):
    rec, g = load_case("hd128")
    hp = synth.HParams(*rec["hparams"])
    tensors = list(synth.synth_model(rec["seed"], hp))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "m.bin")
        synth.write_ggjt(path, hp, tensors, f16=f16)
        vocab, model = L.LoadModel(path)
    assert model.hp == hp and len(vocab) == hp.vocab and vocab[300] == b"w300;"
    for name, arr in tensors:
        want = arr.astype(np.float16).astype(np.float32) if (f16 and arr.ndim == 2) else arr
        np.testing.assert_array_equal(model.get_tensor(name, arr.shape), want, err_msg=name)
    if not f16:
        lctx = L.NewContext(model, rec["context"])
        lg = L.Eval(lctx, g["prompt_ids"], 0)
        ref = g["prompt_all_logits"][-1]
        assert np.abs(lg - ref).max() <= 1e-3 * np.abs(ref).max()


def test_load_ggjt_q8_and_stage_range(L, synth):
    hp = synth.HParams(320, 128, 32, 4, 3)
    tensors = list(synth.synth_model(4, hp))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "m.bin")
        synth.write_ggjt(path, hp, tensors)
        _, mq = L.LoadModel(path, weight_type=L.LB_TYPE_Q8_0)
        _, stage = L.LoadModel(path, layer_begin=1, layer_end=3)
    for name, arr in tensors:
        if synth.is_q8_matrix(name):
            np.testing.assert_array_equal(mq.get_tensor(name, arr.shape), synth.dequantize_q8(*synth.quantize_q8(arr)))
        else:
            np.testing.assert_array_equal(mq.get_tensor(name, arr.shape), arr)
    np.testing.assert_array_equal(stage.get_tensor("layers.2.feed_forward.w2.weight", (128, hp.ff)), dict(tensors)["layers.2.feed_forward.w2.weight"])
    with pytest.raises(L.LlamaB200Error):
        stage.get_tensor("layers.0.attention.wq.weight", (128, 128))      # not held by this stage
    with pytest.raises(L.LlamaB200Error):
        stage.get_tensor("tok_embeddings.weight", (320, 128))


def test_load_ggjt_rejects_like_the_reference(L, synth):
    hp = synth.HParams(64, 64, 32, 2, 1)
    with tempfile.TemporaryDirectory() as td:
        good = os.path.join(td, "good.bin")
        synth.write_ggjt(good, hp, synth.synth_model(1, hp))
        raw = open(good, "rb").read()
        cases = {
            "magic": struct.pack("<I", 0x12345678) + raw[4:],                    # llama.go:729
            "old": struct.pack("<I", 0x67676d6c) + raw[4:],                      # llama.go:724
            "version": raw[:4] + struct.pack("<I", 2) + raw[8:],                 # llama.go:736
            "truncated": raw[: len(raw) - 100],                                  # llama.go:951-955
            "unknown": raw.replace(b"norm.weight", b"norm.weighx", 1),           # llama.go:906-910
        }
        for what, blob in cases.items():
            p = os.path.join(td, what + ".bin")
            open(p, "wb").write(blob)
            with pytest.raises(L.LlamaB200Error):
                L.LoadModel(p)
        with pytest.raises(L.LlamaB200Error):
            L.LoadModel(os.path.join(td, "missing.bin"))
