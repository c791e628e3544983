#!/usr/bin/env python
"""Design study for DESIGN.md §9 item 3 (CPU only, numpy): can the Q8_0 decode GEMV run on the int8 tensor cores
without changing its results?

Scheme: a Q8 block is 32 weights q_k (int8) with one fp32 scale d.  For the same 32 activations x_k take the power of
two 2^e just above max|x_k| and write x_k / 2^e as ND balanced base-128 digits (each in [-64, 64], an int8):
x_k = 2^e * sum_j dig_jk * 128^-(j+1) + residual.  Then sum_k q_k x_k = 2^e * sum_j 128^-(j+1) * (sum_k q_k dig_jk),
where the inner sums are exact s32 dot products (`mma.sync.m16n8k32.s8`, the digits being columns of B).

Prints the maximum error relative to max|y| against an f64 reference for: the FP32 FMA chain the CUDA-core kernel
uses today, and the digit scheme with 3 and 4 digits.  Result (seed 0, K = 4096, heavy-tailed activations):
    f32 chain 8.9e-07 | 3 digits 1.6e-06 | 4 digits 2.6e-07
i.e. 4 digits (4 of the 8 B columns) are more exact than the chain they would replace."""
import numpy as np


def main():
    rs = np.random.RandomState(0)
    K, M = 4096, 256
    W = rs.randn(M, K).astype(np.float32) / 64
    Wb = W.reshape(M, K // 32, 32)
    d = (np.abs(Wb).max(-1) / 127).astype(np.float32)
    q = np.rint(Wb / np.where(d == 0, 1, d)[..., None]).clip(-127, 127).astype(np.int32)
    x = (rs.randn(K) * np.exp(rs.randn(K))).astype(np.float32)
    xb = x.reshape(K // 32, 32)
    ref = (d[..., None].astype(np.float64) * q).reshape(M, K) @ x.astype(np.float64)
    nrm = np.abs(ref).max()

    acc = np.zeros(M, np.float32)                      # today's kernel: d * sum_4(q x) per 4 weights, FP32
    for b in range(K // 32):
        for g in range(8):
            t = np.zeros(M, np.float32)
            for i in range(4):
                t = (t + q[:, b, g * 4 + i].astype(np.float32) * xb[b, g * 4 + i]).astype(np.float32)
            acc = (acc + d[:, b] * t).astype(np.float32)
    print("f32 chain          max rel err %.2e" % (np.abs(acc - ref).max() / nrm))

    for nd in (3, 4):
        acc = np.zeros(M, np.float32)
        for b in range(K // 32):
            xv = xb[b].astype(np.float64)
            mx = np.abs(xv).max()
            if mx == 0:
                continue
            e = int(np.ceil(np.log2(mx))) + 1          # |x| / 2^e <= 0.5
            r, scale, tot = xv / 2.0 ** e, 1.0, np.zeros(M)
            for _ in range(nd):
                scale *= 128
                dig = np.rint(r * scale)
                assert np.abs(dig).max() <= 64
                r = r - dig / scale
                dj = q[:, b, :].astype(np.int64) @ dig.astype(np.int64)
                assert np.abs(dj).max() < 2 ** 31      # fits the s32 accumulator: 32 * 127 * 64 = 2.6e5
                tot += dj / scale
            acc = (acc + (d[:, b].astype(np.float64) * 2.0 ** e * tot).astype(np.float32)).astype(np.float32)
        print("%d base-128 digits  max rel err %.2e" % (nd, np.abs(acc - ref).max() / nrm))


if __name__ == "__main__":
    main()
