"""Q8_0 block-quantised MulMat path (BASELINE config 3).  No reference implementation exists; the
parity target is the reference FP32 path — here the oracle pinned to the
reference binary — run on the DEQUANTISED weights d*q."""
import numpy as np
import pytest

from conftest import load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import llama
    return llama


def test_device_quantiser_matches_numpy_definition(L, synth):
    hp = synth.HParams(320, 128, 32, 4, 2)
    model = L.Model(hp, weight_type=L.LB_TYPE_Q8_0)
    tensors = dict(synth.synth_model(3, hp))
    # edge cases inside one matrix: an all-zero block, a block with a single huge value, exact .5 ties
    w = tensors["layers.0.attention.wq.weight"].copy()
    w[0, :32] = 0.0
    w[1, :32] = 0.0; w[1, 5] = 1000.0
    w[2, :32] = np.arange(32, dtype=np.float32) * 0.5 - 8.0; w[2, 31] = 63.5
    tensors["layers.0.attention.wq.weight"] = w
    model.load(tensors.items())
    for name, arr in tensors.items():
        got = model.get_tensor(name, arr.shape)
        if synth.is_q8_matrix(name):
            q, d = synth.quantize_q8(arr)
            np.testing.assert_array_equal(got, synth.dequantize_q8(q, d), err_msg=name)
            assert np.abs(got - arr).max() <= np.abs(arr).max() / 127.0 * 0.5001 + 1e-12
        else:
            np.testing.assert_array_equal(got, arr, err_msg=name)   # vectors and embeddings stay FP32


@pytest.mark.parametrize("case", ["tiny", "hd128"])
def test_q8_eval_matches_oracle_on_dequantised_weights(L, synth, oracle, case):
    rec, g = load_case(case)
    hp = synth.HParams(*rec["hparams"])
    tensors = list(synth.synth_model(rec["seed"], hp))
    model = L.Model(hp, weight_type=L.LB_TYPE_Q8_0).load(tensors)
    deq = [(n, synth.dequantize_q8(*synth.quantize_q8(a)) if synth.is_q8_matrix(n) else a) for n, a in tensors]
    om = oracle.OracleModel(hp).load(deq)
    oc = oracle.OracleContext(om, rec["context"])
    lctx = L.NewContext(model, rec["context"])
    ids = g["prompt_ids"]                       # 30-33 tokens: the fused-dequant GEMM path
    got = L.EvalAllLogits(lctx, ids, 0)
    _, ref = oc.eval(ids, 0, all_logits=True)
    assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max()
    past = len(ids)
    for tok in g["gen_ids"][:6]:                # GEMV path (first step eager, then CUDA-graph replay)
        got = L.Eval(lctx, [int(tok)], past).copy()
        ref = oc.eval([int(tok)], past)
        past += 1
        assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max()
        assert np.all(np.abs(got - ref) <= 1e-3 * np.abs(ref) + 1e-3 * np.abs(ref).max())
    # quantisation itself moves the logits by far more than the kernel error: the test is meaningful
    fp32_ref = g["step_logits"][6]
    assert np.abs(ref - fp32_ref).max() > 10 * np.abs(got - ref).max()
    # a short prompt (N <= 8) goes through the multi-column GEMV
    l2 = L.NewContext(model, rec["context"])
    o2 = oracle.OracleContext(om, rec["context"])
    got = L.Eval(l2, ids[:5], 0).copy()
    ref = o2.eval(ids[:5], 0)
    assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max()


def test_q8_7b_shaped_layers_against_oracle(L, synth, oracle):
    hp = synth.HParams(32000, 4096, 256, 32, 2)
    model = L.Model(hp, weight_type=L.LB_TYPE_Q8_0).init_random(0)
    om = oracle.OracleModel(hp)
    for name, _tid, shape, _m, _s in synth.tensor_table(hp):
        om.set_tensor(name, model.get_tensor(name, shape))      # dequantised d*q from the device
    oc = oracle.OracleContext(om, 32)
    lctx = L.NewContext(model, 32)
    ids = [1, 35, 35, 107, 104, 111, 31999, 0, 2024]
    got = L.Eval(lctx, ids, 0).copy()
    ref = oc.eval(ids, 0)
    e0 = np.abs(got - ref).max() / np.abs(ref).max()
    assert e0 <= 1e-3
    got = L.Eval(lctx, [17], len(ids)).copy()
    ref = oc.eval([17], len(ids))
    e1 = np.abs(got - ref).max() / np.abs(ref).max()
    assert e1 <= 1e-3
    print(f"Q8 7B-shaped: prefill rel err {e0:.3e}, decode rel err {e1:.3e}")
    full = L.Model(synth.LLAMA_7B, weight_type=L.LB_TYPE_Q8_0)
    assert full.weight_bytes_per_token == 6607077376 // 32 * 36 + (65 * 4096 + 4096) * 4   # 7.43 GB


def test_q8_rejects_unsupported_shapes(L, synth):
    with pytest.raises(L.LlamaB200Error):
        L.Model(synth.HParams(64, 48, 16, 2, 1), weight_type=L.LB_TYPE_Q8_0)   # dim not a multiple of 32
