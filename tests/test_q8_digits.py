"""CPU restatement (numpy, float32 step by step) of the arithmetic of the Q8 ring megakernel's MulMat
(llama.go_b200/csrc/kernels_ring_q8.cu: emit4 + consume), checked against an f64 reference of the parity target —
the FP32 path on the dequantised weights d*q (tests/test_gpu_q8.py compares the GPU with the oracle on exactly that):

* a block of 32 activations is written as four balanced base-128 digit planes relative to the block's power-of-two
  scale s (exponent bits of the block maximum + 1):  x = s * (d0/2^6 + d1/2^13 + d2/2^20 + d3/2^27) + r, |r| <= s * 2^-28,
  every digit an int8 in [-64, 64];
* the int8 x int8 dot products over a block are exact in s32 (|sum| <= 32 * 127 * 64 < 2^22);
* per (row, block):  acc += (c0 * 2^-6 + c1 * 2^-13  [+ lane t = 1: c2 * 2^-20 + c3 * 2^-27]) * (d_w * s), FP32.
The result must be as close to the f64 reference as a plain FP32 FMA chain over the dequantised weights."""
import numpy as np
import pytest

F = np.float32


def digit_planes(xb):
    """emit4: xb [nblk][32] float32 -> digits [nblk][4][32] int32, scales [nblk] float32"""
    m = np.abs(xb).max(-1)
    E = (m.view(np.uint32) >> 23) & 0xFF
    s = np.where(E > 0, ((E + 1).astype(np.uint32) << 23).view(np.float32), F(1.0)).astype(F)
    inv = np.where(E > 0, ((253 - E).astype(np.uint32) << 23).view(np.float32), F(0.0)).astype(F)
    r = (xb * inv[:, None]).astype(F) * F(64.0)
    digs = []
    for j in range(4):
        d = np.rint(r).astype(np.int32)                    # __float2int_rn: round to nearest even
        digs.append(d)
        r = (r - d.astype(F)).astype(F)
        if j < 3:
            r = (r * F(128.0)).astype(F)
    return np.stack(digs, 1), s


@pytest.mark.parametrize("seed,heavy", [(0, False), (1, True), (2, True)])
def test_digit_planes_reconstruct_the_activations(seed, heavy):
    rs = np.random.RandomState(seed)
    x = rs.randn(64, 32)
    if heavy:
        x *= np.exp(2 * rs.randn(64, 32))
    x[3] = 0.0                                              # an all-zero block is sent as zeros
    x[5, 1:] = 0.0
    xb = x.astype(F)
    dig, s = digit_planes(xb)
    assert np.abs(dig).max() <= 64
    w = np.array([2.0 ** -6, 2.0 ** -13, 2.0 ** -20, 2.0 ** -27])
    rec = s[:, None].astype(np.float64) * np.einsum("bjk,j->bk", dig.astype(np.float64), w)
    assert np.all(np.abs(rec - xb.astype(np.float64)) <= s[:, None].astype(np.float64) * 2.0 ** -28 * 1.0000001)
    assert np.all(dig[3] == 0) and s[3] == 1.0


def quantize(W):
    Wb = W.reshape(W.shape[0], -1, 32)
    d = (np.abs(Wb).max(-1) / 127).astype(F)
    q = np.rint(Wb / np.where(d == 0, 1, d)[..., None]).clip(-127, 127).astype(np.int32)
    return q, d


@pytest.mark.parametrize("K", [256, 4096, 11008])
def test_int8_digit_mulmat_matches_f64_of_the_dequantised_weights(K):
    rs = np.random.RandomState(K)
    M = 64
    q, d = quantize((rs.randn(M, K) / 64).astype(F))
    x = (rs.randn(K) * np.exp(rs.randn(K))).astype(F)
    ref = (d[..., None].astype(np.float64) * q).reshape(M, K) @ x.astype(np.float64)
    nrm = np.abs(ref).max()

    dig, s = digit_planes(x.reshape(-1, 32))
    # the kernel: lanes t = 0 / t = 1 of a row carry digit pairs (0, 1) / (2, 3); each accumulates
    # fma(va, d_w * s, acc) per block in FP32; the two lanes' sums are added once per tile; 16 warps each own 2 of every
    # 32 blocks and are summed in warp order
    w = [F(2.0 ** -6), F(2.0 ** -13), F(2.0 ** -20), F(2.0 ** -27)]
    nblk = K // 32
    part = np.zeros((16, 2, M), F)
    for b in range(nblk):
        c = np.einsum("mk,jk->jm", q[:, b, :].astype(np.int64), dig[b].astype(np.int64))      # exact s32 products
        assert np.abs(c).max() < 2 ** 22
        scale = (d[:, b] * s[b]).astype(F)
        warp = (b % 32) // 2
        for t in range(2):
            va = (c[2 * t + 1].astype(F) * w[2 * t + 1] + (c[2 * t].astype(F) * w[2 * t]).astype(F)).astype(F)
            part[warp, t] = (va * scale + part[warp, t]).astype(F)
    lanes = (part[:, 0] + part[:, 1]).astype(F)
    got = np.zeros(M, F)
    for wv in range(16):
        got = (got + lanes[wv]).astype(F)

    chain = np.zeros(M, F)                                   # a plain FP32 FMA chain over the dequantised weights
    deq = (d[..., None] * q.astype(F)).astype(F).reshape(M, K)
    for k in range(K):
        chain = (deq[:, k] * x[k] + chain).astype(F)

    err_digits = np.abs(got - ref).max() / nrm
    err_chain = np.abs(chain - ref).max() / nrm
    assert err_digits <= 2e-6, err_digits                    # three orders below the 1e-3 budget
    assert err_digits <= 2 * err_chain + 2e-7, (err_digits, err_chain)
