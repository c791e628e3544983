"""SampleTopPTopK (pkg/llama/llama.go:455-707), the context-swap rule and the generate loop of server.Do
(pkg/server/server.go:127-237).

CPU: the oracle's restatement of the loop reproduces the reference binary's stream PAST the context (three swaps;
tests/golden/refbin_swap.json, tools/gen_golden.py swap), lb_context_swap (pure host) equals the oracle's rule.
GPU: the device sampler's candidate set (ids + probabilities after the top-k and top-p cuts) equals
oracle.sample_candidates on the same logits; lb_generate reproduces the reference binary's stream through the swaps."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def load_swap():
    with open(os.path.join(GOLDEN, "refbin_swap.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("mode", ["scalar", "avx"])
def test_oracle_generate_loop_reproduces_reference_binary_through_context_swaps(oracle, synth, mode):
    from oracle import refbin
    rec = load_swap()
    hp = synth.HParams(*rec["hparams"])
    oracle.set_dot_mode(mode == "avx")
    try:
        m = oracle.OracleModel(hp).load(synth.synth_model(rec["seed"], hp))
        toks = oracle.generate_stream(oracle.OracleContext(m, rec["context"]), rec["prompt_ids"], rec["predict"], rec["context"])
    finally:
        oracle.set_dot_mode(False)
    assert toks == rec["oracle_tokens"]
    vocab = synth.byte_vocab(hp.vocab)
    assert refbin.same_stream(bytes.fromhex(rec["runs"][mode]["text_hex"]), refbin.expected_text(vocab, rec["prompt_ids"], toks))
    assert len(rec["prompt_ids"]) + rec["predict"] > rec["context"] + 20      # the run really went past the context
    assert rec["runs"][mode]["evals"] == rec["predict"]


def test_context_swap_rule_matches_oracle(oracle):
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import llama
    rs = np.random.RandomState(0)
    for ctx, keep, past, n_embd in [(40, 0, 40, 1), (40, 0, 39, 1), (40, 4, 40, 1), (64, 8, 60, 9), (32, 0, 32, 3), (16, 0, 5, 2)]:
        hist = rs.randint(0, 500, size=ctx).astype(np.uint32)
        embd = rs.randint(0, 500, size=n_embd).astype(np.uint32)
        ref_past, ref_embd = oracle.context_swap(ctx, keep, hist.tolist(), past, embd.tolist())
        got_past, got_embd = llama.ContextSwap(ctx, keep, hist, past, embd)
        assert (got_past, got_embd) == (ref_past, ref_embd)
        if past + n_embd > ctx:       # server.go:166-171: pastCount = keep; (past - keep) / 2 most recent ids in front
            assert got_past == keep and len(got_embd) == (past - keep) // 2 + n_embd
            assert got_embd[:len(got_embd) - n_embd] == hist[ctx - (past - keep) // 2:].tolist()
    with pytest.raises(llama.LlamaB200Error):
        llama.ContextSwap(40, 50, np.zeros(40, np.uint32), 40, [1])     # keep > pastCount


def test_oracle_sampler_shapes_and_limits(oracle):
    rs = np.random.RandomState(1)
    lg = rs.standard_normal(512).astype(np.float32) * 3
    ids, probs = oracle.sample_candidates(lg, [1, 2, 3], 40, 0.95, 0.8, 1.1)
    assert 1 <= len(ids) <= 40 and abs(float(probs.sum()) - 1.0) < 1e-5 and np.all(np.diff(probs) <= 0)
    ids1, probs1 = oracle.sample_candidates(lg, [1, 2, 3], 40, 0.95, 1e-6, 1.1)      # temp -> 0: one candidate
    assert len(ids1) == 1 and probs1[0] == 1.0
    full, pfull = oracle.sample_candidates(lg, [], 512, 1.0, 1.0, 1.0)
    assert len(full) == 512 and full[0] == int(np.argmax(lg))
    assert oracle.sample_pick(ids1, probs1, 5) == int(ids1[0])


@pytest.mark.gpu
@pytest.mark.parametrize("top_k,top_p,temp,penalty", [(40, 0.95, 0.8, 1.1), (40, 1.0, 0.8, 1.1), (1, 0.95, 0.8, 1.1), (200, 0.5, 1.3, 1.0),
                                                      (40, 0.95, 1e-6, 1.1), (512, 0.999, 2.0, 1.3)])
def test_device_sampler_candidate_set_matches_oracle(oracle, synth, top_k, top_p, temp, penalty):
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import llama
    rec = load_swap()
    hp = synth.HParams(*rec["hparams"])
    model = llama.Model(hp).load(synth.synth_model(rec["seed"], hp))
    lctx = llama.NewContext(model, 64)
    logits = llama.Eval(lctx, rec["prompt_ids"], 0).copy()
    last_n = [0] * 30 + list(rec["prompt_ids"])
    for seed in (0, 1, 12345):
        tok, ids, probs = llama.SampleTopPTopK(lctx, last_n, top_k, top_p, temp, penalty, seed)
        rids, rprobs = oracle.sample_candidates(logits, last_n, top_k, top_p, temp, penalty)
        np.testing.assert_array_equal(ids, rids)                        # index work: exact
        np.testing.assert_allclose(probs, rprobs, rtol=3e-7, atol=0)    # f64 exp of two libms, then identical FP32 steps
        assert tok == oracle.sample_pick(rids, rprobs, seed) or not np.array_equal(probs, rprobs)
        assert tok in ids.tolist()
    with pytest.raises(llama.LlamaB200Error):
        llama.SampleTopPTopK(lctx, last_n, hp.vocab + 1, 0.9, 0.8, 1.1)   # the reference's slice logitsID[:topK] would panic
    with pytest.raises(llama.LlamaB200Error):
        llama.SampleTopPTopK(lctx, last_n, 0, 0.9, 0.8, 1.1)


@pytest.mark.gpu
def test_generate_through_context_swaps_reproduces_reference_binary(synth):
    """lb_generate = server.Do's loop on the engine (device sampler at the reference's temp 1e-6 greedy limit, context
    swap with keep 0): 40 tokens from a 30-token prompt in a 40-token context must be the reference binary's stream."""
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import llama
    from oracle import refbin
    rec = load_swap()
    hp = synth.HParams(*rec["hparams"])
    model = llama.Model(hp).load(synth.synth_model(rec["seed"], hp))
    lctx = llama.NewContext(model, rec["context"])
    toks = llama.Generate(lctx, rec["prompt_ids"], rec["predict"], 40, 0.95, 1e-6, 1.10, 0, rec["context"], seed=3)
    assert toks == rec["oracle_tokens"]
    vocab = synth.byte_vocab(hp.vocab)
    for mode in ("scalar", "avx"):
        assert refbin.same_stream(bytes.fromhex(rec["runs"][mode]["text_hex"]), refbin.expected_text(vocab, rec["prompt_ids"], toks))
    # without the swap the same request is rejected by Eval (pastCount + N > context), as lb_generate_greedy documents
    with pytest.raises(llama.LlamaB200Error):
        llama.GenerateGreedy(llama.NewContext(model, rec["context"]), rec["prompt_ids"], rec["predict"])
