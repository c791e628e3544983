"""tcgen05/TMEM/TMA prefill GEMM (csrc/kernels_tc.cu, 3xTF32 split) against the CPU oracle's
ComputeForwardMulMatFP32.  FP32 reference semantics: the tensor-core path must stay FP32-class
(products accurate to ~2^-20), far inside the 1e-3 logits budget."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ml():
    import llama_go_b200  # noqa: F401
    from llama_go_b200 import ml as M
    return M


def run_mul_mat(ml, w, x):
    ctx = ml.NewContext()
    M, K = w.shape
    N = x.shape[0]
    a = ml.NewTensor2D(ctx, ml.TYPE_F32, K, M, w)
    b = ml.NewTensor2D(ctx, ml.TYPE_F32, K, N, x)
    c = ml.MulMat(ctx, a, b)
    g = ml.Graph()
    ml.BuildForwardExpand(g, c)
    ml.GraphCompute(ctx, g)
    out = c.numpy().reshape(N, M)
    ctx.ReleaseContext()
    return out


@pytest.mark.parametrize("M,K,N", [(128, 32, 9), (128, 64, 16), (192, 64, 30), (704, 256, 33), (300, 128, 200),
                                   (64, 96, 129), (1024, 1024, 128), (4096, 4096, 64), (4096, 11008, 40)])
def test_tensor_core_gemm_matches_oracle(ml, oracle, M, K, N):
    rng = np.random.default_rng(M * 7 + K * 3 + N)
    w = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    x = rng.standard_normal((N, K)).astype(np.float32)
    got = run_mul_mat(ml, w, x)
    ref = oracle.op_mul_mat_2d(w, x)
    # exact product sum in float64 as the yardstick for both
    exact = x.astype(np.float64) @ w.astype(np.float64).T
    e_gpu = np.abs(got - exact).max() / np.abs(exact).max()
    e_ref = np.abs(ref - exact).max() / np.abs(exact).max()
    print(f"[{M}x{K}x{N}] tcgen05 3xTF32 err {e_gpu:.2e}   reference FP32 loop err {e_ref:.2e}")
    assert np.isfinite(got).all()
    assert e_gpu < 2e-5, "tensor-core GEMM lost FP32-class accuracy"
    assert np.abs(got - ref).max() <= 3e-5 * np.abs(ref).max()


def test_tensor_core_gemm_handles_special_values(ml):
    w = np.zeros((128, 64), np.float32)
    x = np.zeros((16, 64), np.float32)
    w[3, 5] = 1.0; x[2, 5] = 3.0
    w[7, :] = 1e-30; x[4, :] = 1e-30          # products underflow to zero cleanly
    w[9, 0] = np.float32(1.0 + 2.0 ** -20); x[1, 0] = 1.0   # FP32 value that TF32 cannot hold: needs the lo term
    got = run_mul_mat(ml, w, x)
    assert got[2, 3] == 3.0
    assert got[1, 9] == np.float32(1.0 + 2.0 ** -20)
    exact = x.astype(np.float64) @ w.astype(np.float64).T     # tiny cross terms (1e-30, 3e-30) are legitimate, 1e-60 underflows
    np.testing.assert_allclose(got, exact.astype(np.float32), rtol=2e-6, atol=0)
    assert got[4, 7] == 0.0
