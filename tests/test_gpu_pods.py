"""Pod batching: B independent sequences ("pods", pkg/server/server.go:84-106) decoded in one
pass over the weights must give every pod exactly what its own llama.Eval gives — each pod keeps its own
KV cache and position."""
import numpy as np
import pytest

from conftest import load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import llama
    return llama


@pytest.mark.parametrize("B", [1, 3, 8])
def test_pod_batch_matches_individual_evals_and_golden(L, synth, B):
    rec, g = load_case("hd128")
    hp = synth.HParams(*rec["hparams"])
    model = L.Model(hp).load(synth.synth_model(rec["seed"], hp))
    ids = [int(t) for t in g["prompt_ids"]]
    gen = [int(t) for t in g["gen_ids"][:-1]]
    # pods at DIFFERENT positions: pod b has consumed the prompt plus b generated tokens
    pods, solo = [], []
    for b in range(B):
        pc, sc = L.NewContext(model, rec["context"]), L.NewContext(model, rec["context"])
        for c in (pc, sc):
            L.Eval(c, ids, 0)
            for i in range(b):
                L.Eval(c, [gen[i]], len(ids) + i)
        pods.append(pc); solo.append(sc)
    batch = L.PodBatch(pods)
    for step in range(4):
        toks = [gen[b + step] for b in range(B)]
        pasts = [len(ids) + b + step for b in range(B)]
        got = batch.Eval(toks, pasts)
        for b in range(B):
            ref_solo = L.Eval(solo[b], [toks[b]], pasts[b])
            ref_gold = g["step_logits"][b + step + 1]
            assert np.abs(got[b] - ref_solo).max() <= 2e-5 * np.abs(ref_solo).max()
            assert np.abs(got[b] - ref_gold).max() <= 1e-3 * np.abs(ref_gold).max()
    # the pods' own caches were written: continuing one pod on its own gives the golden logits
    b = B - 1
    lg = L.Eval(pods[b], [gen[b + 4]], len(ids) + b + 4)
    ref = g["step_logits"][b + 5]
    assert np.abs(lg - ref).max() <= 1e-3 * np.abs(ref).max()


def test_pod_batch_resident_and_checks(L, synth):
    rec, g = load_case("tiny")
    hp = synth.HParams(*rec["hparams"])
    model = L.Model(hp).load(synth.synth_model(rec["seed"], hp))
    ids = [int(t) for t in g["prompt_ids"]]
    gen = np.asarray(g["gen_ids"][:-1], np.uint32)
    pods = [L.NewContext(model, rec["context"]) for _ in range(4)]
    for c in pods:
        L.Eval(c, ids, 0)
    batch = L.PodBatch(pods)
    ms = batch.DecodeResident(np.stack([gen] * 4), [len(ids)] * 4)
    assert ms > 0
    ref = g["step_logits"][len(gen)]
    for row in batch.ReadLogits():
        assert np.abs(row - ref).max() <= 1e-3 * np.abs(ref).max()
    with pytest.raises(L.LlamaB200Error):
        L.PodBatch([pods[0], pods[0]])
    other = L.NewContext(model, rec["context"] * 2)
    with pytest.raises(L.LlamaB200Error):
        L.PodBatch([pods[0], other])
    with pytest.raises(L.LlamaB200Error):
        batch.Eval([1, 2, 3, 4], [rec["context"]] * 4)        # position outside the context
    with pytest.raises(L.LlamaB200Error):
        L.PodBatch([L.NewContext(model, 64) for _ in range(9)])


@pytest.mark.parametrize("B", [8, 2])
def test_pod_batch_7b_shaped_at_T400_matches_solo_decodes(L, synth, B):
    """Exact LLaMA-7B layer shapes (2 layers, vocab cut to 2048), context 512, pods at positions 400 + 3b: the
    pod-batch megakernel (B-column MulMat on the tensor cores, 3xTF32; w2 walks K = 11008 in 4 streamed passes;
    B = 8 -> one attention item per (pod, head), B = 2 -> 4 splits + merge) against each pod's own single-sequence
    decode, which tests/test_gpu_longctx.py checks against the oracle at the same T."""
    hp = synth.HParams(2048, 4096, 256, 32, 2)
    model = L.Model(hp).init_random(0)
    rs = np.random.RandomState(3)
    pods, solo, pasts = [], [], []
    for b in range(B):
        n = 400 + 3 * b
        ids = rs.randint(3, hp.vocab, size=n).astype(np.uint32)
        pc, sc = L.NewContext(model, 512), L.NewContext(model, 512)
        L.Eval(pc, ids, 0)
        L.Eval(sc, ids, 0)
        pods.append(pc); solo.append(sc); pasts.append(n)
    batch = L.PodBatch(pods)
    worst = 0.0
    for step in range(3):
        toks = rs.randint(3, hp.vocab, size=B).astype(np.uint32)
        got = batch.Eval(toks, [p + step for p in pasts])
        for b in range(B):
            ref = L.Eval(solo[b], [int(toks[b])], pasts[b] + step)
            err = np.abs(got[b] - ref).max() / np.abs(ref).max()
            worst = max(worst, err)
            assert err <= 2e-5, f"pod {b} step {step}: rel err {err:.3e}"
            assert int(np.argmax(got[b])) == int(np.argmax(ref))
    # the pods' caches hold what the batch wrote: K/V rows of the last step equal the solo contexts'
    for b in (0, B - 1):
        kp, vp = pods[b].kv(1, pasts[b] + 2, 1)
        ks, vs = solo[b].kv(1, pasts[b] + 2, 1)
        np.testing.assert_allclose(kp, ks, rtol=0, atol=2e-5 * np.abs(ks).max())
        np.testing.assert_allclose(vp, vs, rtol=0, atol=2e-5 * np.abs(vs).max())
    print(f"[pods B={B}, 7B-shaped, T~400] worst rel err vs solo {worst:.3e}")
