"""ml.Tokenize on the host (csrc/tokenizer.cpp).  Pinned to the reference through a golden
case whose vocab HAS merges: the reference binary's generated stream for the prompt can only equal the
oracle's stream if both started from the same prompt ids."""
import pytest

from conftest import load_case


@pytest.fixture(scope="module")
def ml():
    import __graft_entry__ as g
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import _capi, ml as M
    import os
    if not os.path.exists(_capi.LIB_PATH):
        g.build()
    return M


def test_tokenizer_matches_reference_binary_through_the_merges_fixture(ml, synth, oracle):
    from oracle import refbin
    rec, _ = load_case("merges")
    hp = synth.HParams(*rec["hparams"])
    vocab, scores = synth.merge_vocab(hp.vocab)
    # main.go:129 and server.go:120 each prepend one space to the prompt; Tokenize adds BOS (ml.go:2767)
    ids = ml.Tokenize(ml.Vocab(vocab, scores), b"  " + rec["prompt"].encode(), True)
    assert ids == rec["prompt_ids"]
    assert ids[:8] == [1, 35, 35, 6, 35, 12, 35, 14]          # BOS, ' ', ' ', hello, ' ', world, ' ', the
    for mode, avx in (("scalar", False), ("avx", True)):
        oracle.set_dot_mode(avx)
        try:
            m = oracle.OracleModel(hp).load(synth.synth_model(rec["seed"], hp))
            c = oracle.OracleContext(m, rec["context"])
            toks = oracle.greedy_stream(c, ids, rec["predict"], rec["context"])
        finally:
            oracle.set_dot_mode(False)
        assert toks == rec["oracle_tokens"]
        assert refbin.same_stream(bytes.fromhex(rec["runs"][mode]["text_hex"]), refbin.expected_text(vocab, ids, toks))


def test_tokenizer_rules(ml):
    toks = [b"<unk>", b"", b"", b"ab", b"bc", b"abc", b"a", b"b", b"c", b"\xc3\xa9", b"aa"]
    sc = [0, 0, 0, -1.0, -1.0, -0.5, -9, -9, -9, -2.0, -1.0]
    v = ml.Vocab(toks, sc)
    assert ml.Tokenize(v, b"", True) == [1] and ml.Tokenize(v, b"", False) == []
    assert ml.Tokenize(v, b"abc", False) == [5]                  # ab (-1) merges first (leftmost of the tie with bc), then ab+c -> abc
    assert ml.Tokenize(v, b"bca", False) == [4, 6]                # bc, a
    assert ml.Tokenize(v, b"aaa", False) == [10, 6]               # ties: the leftmost pair merges first -> aa, a
    assert ml.Tokenize(v, b"xyz", False) == [ord("x") + 3, ord("y") + 3, ord("z") + 3]   # byte fallback id = byte + 3 (ml.go:2831)
    assert ml.Tokenize(v, "é".encode(), False) == [9]             # one 2-byte UTF-8 character = one symbol
    assert ml.Tokenize(v, "ü".encode(), False) == [0xC3 + 3, 0xBC + 3]
    assert ml.Tokenize(v, b"a\xe2\x82", False) == [6, 0xE2 + 3, 0x82 + 3]   # truncated multi-byte tail is clamped (ml.go:2777)
    assert ml.Tokenize(v, b"ab", True) == [1, 3]
    # Go computes `symbol.Text[j] + 3` in byte arithmetic (ml.go:2831): 0xFD, 0xFE, 0xFF wrap to ids 0, 1, 2
    assert ml.Tokenize(v, b"\xfd\xfe\xff", False) == [0, 1, 2]
    assert ml.Tokenize(v, b"a\xfc\xfd", False) == [6, 0xFF, 0]
    assert ml.Tokenize(v, b"\xff", True) == [1, 2]
