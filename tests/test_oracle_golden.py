"""Oracle regression against the committed golden logits, and oracle self-consistency."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_case


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_matches_golden_logits(oracle, synth, case):
    rec, g = load_case(case)
    hp = synth.HParams(*rec["hparams"])
    m = oracle.OracleModel(hp).load(synth.synth_model(rec["seed"], hp))
    c = oracle.OracleContext(m, rec["context"])
    ids = g["prompt_ids"]
    last, allrows, hid = c.eval(ids, 0, all_logits=True, hidden=True)
    # same code, same inputs: only libm's last-ulp behaviour may differ between hosts
    np.testing.assert_allclose(allrows, g["prompt_all_logits"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(hid, g["prompt_hidden"], rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(last, allrows[-1])
    k, v = c.kv()
    np.testing.assert_allclose(k[:, :len(ids)], g["k_after_prompt"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(v[:, :len(ids)], g["v_after_prompt"], rtol=1e-5, atol=1e-6)
    past = len(ids)
    for i, tok in enumerate(g["gen_ids"][:-1]):
        lg = c.eval([int(tok)], past)
        past += 1
        np.testing.assert_allclose(lg, g["step_logits"][i + 1], rtol=1e-5, atol=1e-5)


def test_prefill_equals_token_by_token(oracle, synth):
    """Evaluating N tokens at once or one at a time fills the same cache and gives the same logits
    (up to dot-order-free FP32 identity: every op is per-row, so this is exact)."""
    hp = synth.HParams(256, 64, 32, 2, 2)
    m = oracle.OracleModel(hp).load(synth.synth_model(3, hp))
    ids = [1, 17, 200, 45, 99, 3, 250, 8, 77]
    a = oracle.OracleContext(m, 32)
    la = a.eval(ids, 0)
    b = oracle.OracleContext(m, 32)
    for i, t in enumerate(ids):
        lb = b.eval([t], i)
    np.testing.assert_array_equal(la, lb)
    np.testing.assert_array_equal(a.kv()[0], b.kv()[0])


def test_eval_rejects_bad_input(oracle, synth):
    hp = synth.HParams(64, 32, 32, 2, 1)
    m = oracle.OracleModel(hp).load(synth.synth_model(1, hp))
    c = oracle.OracleContext(m, 8)
    with pytest.raises(ValueError):
        c.eval([], 0)
    with pytest.raises(ValueError):
        c.eval([1] * 9, 0)       # past + N > ctx
    with pytest.raises(ValueError):
        c.eval([64], 0)          # token id out of range
