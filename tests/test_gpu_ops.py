"""GPU parity, op by op: every pkg/ml op of the C-ABI mirror against the CPU oracle's restatement
of the same ComputeForward* kernel, on the same seeded inputs.  Bit-exact where the op has no
reduction or transcendental; otherwise within the stated tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ml():
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import ml as M
    return M


@pytest.fixture()
def ctx(ml):
    c = ml.NewContext()
    yield c
    c.ReleaseContext()


def run(ml, ctx, t):
    g = ml.Graph()
    ml.BuildForwardExpand(g, t)
    ml.GraphCompute(ctx, g)
    return t.numpy()


rng = np.random.default_rng(1234)


def randn(*shape):
    return rng.standard_normal(shape).astype(np.float32)


def test_get_rows(ml, ctx, oracle):
    table = randn(50, 64)
    ids = np.array([3, 49, 0, 7, 7], np.float32)     # ids travel as float32 (llama.go:239-242)
    a = ml.NewTensor2D(ctx, ml.TYPE_F32, 64, 50, table)
    b = ml.NewTensor1D(ctx, ml.TYPE_F32, 5, ids)
    out = run(ml, ctx, ml.GetRows(ctx, a, b)).reshape(5, 64)
    np.testing.assert_array_equal(out, oracle.op_get_rows(table, ids))


@pytest.mark.parametrize("nc,nr", [(64, 1), (4096, 3), (100, 7)])
def test_rms_norm(ml, ctx, oracle, nc, nr):
    x = randn(nr, nc) * 3
    a = ml.NewTensor2D(ctx, ml.TYPE_F32, nc, nr, x)
    out = run(ml, ctx, ml.RMSNorm(ctx, a)).reshape(nr, nc)
    ref = oracle.op_rms_norm(x)
    # f64 accumulation on both sides; only the association order differs -> scale equal to ~1 ulp
    np.testing.assert_allclose(out, ref, rtol=2e-7, atol=0)


def test_repeat_mul_add_silu_scale(ml, ctx, oracle):
    w = randn(128)
    x = randn(5, 128)
    tw = ml.NewTensor1D(ctx, ml.TYPE_F32, 128, w)
    tx = ml.NewTensor2D(ctx, ml.TYPE_F32, 128, 5, x)
    rep = ml.Repeat(ctx, tw, tx)
    np.testing.assert_array_equal(run(ml, ctx, rep).reshape(5, 128), oracle.op_repeat(w, 5))
    mul = ml.Mul(ctx, rep, tx)
    np.testing.assert_array_equal(run(ml, ctx, mul).reshape(5, 128), oracle.op_mul(oracle.op_repeat(w, 5), x))
    y = randn(5, 128)
    ty = ml.NewTensor2D(ctx, ml.TYPE_F32, 128, 5, y)
    np.testing.assert_array_equal(run(ml, ctx, ml.Add(ctx, tx, ty)).reshape(5, 128), oracle.op_add(x, y))
    big = randn(5, 128) * 8
    tb = ml.NewTensor2D(ctx, ml.TYPE_F32, 128, 5, big)
    # f64 exp on both sides: CUDA's and glibc's exp may differ in the last f64 ulp -> <= 1 f32 ulp
    np.testing.assert_allclose(run(ml, ctx, ml.Silu(ctx, tb)).reshape(5, 128), oracle.op_silu(big), rtol=1.2e-7, atol=1e-38)
    sc = ml.Scale(ctx, tx, ml.NewFP32(ctx, 0.0883883461356163))
    np.testing.assert_array_equal(run(ml, ctx, sc).reshape(5, 128), oracle.op_scale(x, np.float32(0.0883883461356163)))
    # Repeat of an equal-shaped tensor returns the tensor itself (ml.go:496-498)
    t1 = ml.NewTensor1D(ctx, ml.TYPE_F32, 128, w)
    t2 = ml.NewTensor1D(ctx, ml.TYPE_F32, 128, w)
    assert ml.Repeat(ctx, t1, t2)._h == t1._h


def test_mul_rejects_different_shapes(ml, ctx):
    a = ml.NewTensor2D(ctx, ml.TYPE_F32, 8, 2)
    b = ml.NewTensor2D(ctx, ml.TYPE_F32, 8, 3)
    with pytest.raises(ml.LlamaB200Error, match="different shapes"):
        ml.Mul(ctx, a, b)


@pytest.mark.parametrize("M,K,N", [(64, 64, 1), (192, 64, 3), (704, 256, 8), (300, 128, 30), (130, 4096, 2), (257, 260, 70)])
def test_mul_mat_2d(ml, ctx, oracle, M, K, N):
    w, x = randn(M, K) / np.sqrt(K), randn(N, K)
    a = ml.NewTensor2D(ctx, ml.TYPE_F32, K, M, w)
    b = ml.NewTensor2D(ctx, ml.TYPE_F32, K, N, x)
    out = run(ml, ctx, ml.MulMat(ctx, a, b)).reshape(N, M)
    ref = oracle.op_mul_mat_2d(w, x)
    # FP32 dot, different association order + FMA: |err| <= ~K * eps * sum|a_i b_i|
    np.testing.assert_allclose(out, ref, rtol=0, atol=2e-6 * np.sqrt(K) * np.abs(ref).max() + 1e-6)


def test_mul_mat_permuted_attention_shapes(ml, ctx, oracle):
    """K·Q with permuted views and V^T·P, exactly the shapes of llama.go:281-325."""
    hd, H, N, T = 32, 4, 3, 11
    kc = randn(T, H, hd)      # cache rows [t][h][d]  == tensor [hd, H, T]
    q = randn(N, H, hd)
    tk = ml.NewTensor3D(ctx, ml.TYPE_F32, hd, H, T, kc)
    tq = ml.NewTensor3D(ctx, ml.TYPE_F32, hd, H, N, q)
    K = ml.Permute(ctx, tk, 0, 2, 1, 3)
    Q = ml.Permute(ctx, tq, 0, 2, 1, 3)
    assert K.NE == [hd, T, H, 1] and K.NB == [4, hd * H * 4, hd * 4, hd * H * T * 4]
    KQ = ml.MulMat(ctx, K, Q)
    assert KQ.NE == [T, N, H, 1]
    out = run(ml, ctx, KQ).reshape(H, N, T)
    ref = oracle.op_mul_mat(kc, (hd, T, H, 1), (1, hd * H, hd, hd * H * T), q, (hd, N, H, 1), (1, hd * H, hd, hd * H * N)).reshape(H, N, T)
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-5)
    # V^T copy: Permute(1,2,0,3) then Copy into [T, hd, H]
    v = randn(T, H, hd)
    tv = ml.NewTensor3D(ctx, ml.TYPE_F32, hd, H, T, v)
    VT = ml.Copy(ctx, ml.Permute(ctx, tv, 1, 2, 0, 3), ml.NewTensor3D(ctx, ml.TYPE_F32, T, hd, H))
    outv = run(ml, ctx, VT).reshape(H, hd, T)
    np.testing.assert_array_equal(outv, np.transpose(v, (1, 2, 0)))


@pytest.mark.parametrize("mode,past", [(0, 0), (0, 5), (1, 0), (1, 7)])
def test_rope(ml, ctx, oracle, mode, past):
    hd, H = 128, 3
    n2 = 4 if mode == 0 else past + 4
    x = randn(n2, H, hd)
    t = ml.NewTensor3D(ctx, ml.TYPE_F32, hd, H, n2, x)
    out = run(ml, ctx, ml.Rope(ctx, t, past, hd, mode)).reshape(n2, H, hd)
    ref = oracle.op_rope(x, past, hd, mode)
    # f64 pow/sin/cos on both sides, cast to f32: equal up to 1 f32 ulp of the larger pair element
    np.testing.assert_allclose(out, ref, rtol=0, atol=2.4e-7 * np.abs(x).max() * 1.5)
    if mode == 1 and past:
        np.testing.assert_array_equal(out[:past], x[:past])  # rows before `past` untouched


def test_diag_mask_and_softmax(ml, ctx, oracle):
    T, N, H, past = 13, 4, 3, 9
    x = randn(H, N, T) * 4
    t = ml.NewTensor3D(ctx, ml.TYPE_F32, T, N, H, x)
    masked = ml.DiagMaskInf(ctx, t, past)
    out = run(ml, ctx, masked).reshape(H, N, T)
    ref = oracle.op_diag_mask_inf(x, past)
    np.testing.assert_array_equal(out, ref)
    sm = run(ml, ctx, ml.SoftMax(ctx, masked)).reshape(H, N, T)
    refsm = oracle.op_soft_max(ref)
    np.testing.assert_allclose(sm, refsm, rtol=1e-6, atol=1e-9)
    assert np.all(sm[:, 0, past + 1:] == 0)


def test_view1d_and_copy_into_cache(ml, ctx):
    cache = ml.NewTensor1D(ctx, ml.TYPE_F32, 64)
    src = ml.NewTensor2D(ctx, ml.TYPE_F32, 8, 2, np.arange(16, dtype=np.float32))
    view = ml.View1D(ctx, cache, 16, 24)
    run(ml, ctx, ml.Copy(ctx, src, view))
    full = cache.numpy().reshape(-1)
    np.testing.assert_array_equal(full[24:40], np.arange(16, dtype=np.float32))
    assert np.all(full[:24] == 0) and np.all(full[40:] == 0)
    with pytest.raises(ml.LlamaB200Error):
        ml.View1D(ctx, cache, 4, 100)


def test_transpose_is_not_computable_like_the_reference(ml, ctx):
    a = ml.NewTensor2D(ctx, ml.TYPE_F32, 4, 3)
    g = ml.Graph()
    ml.BuildForwardExpand(g, ml.Transpose(ctx, a))
    with pytest.raises(ml.LlamaB200Error, match="transpose"):
        ml.GraphCompute(ctx, g)  # ml.go:1664-1667: "[HALT] Please implement"
