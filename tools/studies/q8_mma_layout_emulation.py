#!/usr/bin/env python
"""CPU emulation of the planned `gemv_q8_mma_kernel`: checks the index math that maps the existing
"4-row interleaved" Q8 planes and the activation digits onto `mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32`
fragments, lane by lane, against d*q.x in f64.  Fragment layout (PTX ISA, m16n8k32 .s8; gid = lane / 4, tig = lane % 4):
  A (16x32 row):  a0 = A[gid][tig*4..+3]   a1 = A[gid+8][tig*4..+3]   a2 = A[gid][16+tig*4..+3]   a3 = A[gid+8][16+tig*4..+3]
  B (32x8 col):   b0 = B[tig*4..+3][gid]   b1 = B[16+tig*4..+3][gid]
  C (16x8):       c0,c1 = C[gid][tig*2+{0,1}]   c2,c3 = C[gid+8][tig*2+{0,1}]
"""
import numpy as np


def quantize_interleaved(W):
    M, K = W.shape
    Wb = W.reshape(M, K // 32, 32)
    d = (np.abs(Wb).max(-1) / 127).astype(np.float32)
    q = np.rint(Wb / np.where(d == 0, 1, d)[..., None]).clip(-127, 127).astype(np.int8).reshape(M, K)
    qplane = np.zeros(M * K, np.int8)     # unit (g, k4) at ((g*(K/4) + k4)*16), byte (r%4)*4 + (k%4)
    dplane = np.zeros(M * K // 32, np.float32)   # (g, kb) at ((g*(K/32) + kb)*4), float r%4
    for r in range(M):
        for k in range(K):
            qplane[((r // 4) * (K // 4) + k // 4) * 16 + (r % 4) * 4 + (k % 4)] = q[r, k]
        for kb in range(K // 32):
            dplane[((r // 4) * (K // 32) + kb) * 4 + (r % 4)] = d[r, kb]
    return q, d, qplane, dplane


def digits_of_block(xv, nd=4):
    """-> exponent scale 2^e (float) and nd int8 digit vectors; x = 2^e * sum_j dig_j * 128^-(j+1) + residual."""
    mx = np.abs(xv).max()
    if mx == 0:
        return np.float32(0), np.zeros((nd, 32), np.int8)
    e = int(np.floor(np.log2(mx))) + 2            # ilogb(max) + 2  ->  |x| / 2^e < 0.5
    r = (xv.astype(np.float32) * np.float32(2.0 ** -e)).astype(np.float32)
    digs = np.zeros((nd, 32), np.int8)
    for j in range(nd):
        r = (r * np.float32(128)).astype(np.float32)          # exact
        dj = np.rint(r).astype(np.float32)
        assert np.abs(dj).max() <= 64
        digs[j] = dj.astype(np.int8)
        r = (r - dj).astype(np.float32)                       # exact
    return np.float32(2.0 ** e), digs


def emulate(M=32, K=128, seed=0):
    rs = np.random.RandomState(seed)
    W = (rs.randn(M, K) / 8).astype(np.float32)
    x = (rs.randn(K) * np.exp(rs.randn(K))).astype(np.float32)
    q, d, qplane, dplane = quantize_interleaved(W)
    ref = (d.repeat(32, axis=1).astype(np.float64) * q) @ x.astype(np.float64)
    qwords = qplane.view(np.int32)                # 4 bytes = 4 consecutive k of one row
    nb = K // 32
    # digit kernel: bfrag[b][lane] = (b0, b1) bytes, xs[b]
    bfrag = np.zeros((nb, 32, 8), np.int8)
    xs = np.zeros(nb, np.float32)
    for b in range(nb):
        xs[b], digs = digits_of_block(x[32 * b:32 * b + 32])
        for lane in range(32):
            gid, tig = lane // 4, lane % 4
            if gid < 4:                            # column n = gid holds digit gid; columns 4..7 are zero
                bfrag[b, lane, 0:4] = digs[gid, tig * 4:tig * 4 + 4]
                bfrag[b, lane, 4:8] = digs[gid, 16 + tig * 4:16 + tig * 4 + 4]
    y = np.zeros(M, np.float32)
    wj = [np.float32(128.0 ** -(j + 1)) for j in range(4)]
    for tile in range(M // 16):
        R0 = tile * 16
        acc = np.zeros((32, 2), np.float32)        # per lane: rows gid, gid+8
        for b in range(nb):
            # --- gather fragments exactly as the kernel will address them
            A = np.zeros((16, 32), np.int64); B = np.zeros((32, 8), np.int64)
            for lane in range(32):
                gid, tig = lane // 4, lane % 4
                for half, (drow, dk4) in enumerate(((0, 0), (8, 0), (0, 4), (8, 4))):   # a0..a3
                    row = R0 + gid + drow
                    k4 = 8 * b + tig + dk4
                    word = ((row // 4) * (K // 4) + k4) * 4 + (row % 4)               # index in 32-bit words
                    bytes4 = qwords[word:word + 1].view(np.int8)
                    A[gid + drow, (16 if dk4 else 0) + tig * 4:(16 if dk4 else 0) + tig * 4 + 4] = bytes4
                B[tig * 4:tig * 4 + 4, gid] = bfrag[b, lane, 0:4]
                B[16 + tig * 4:16 + tig * 4 + 4, gid] = bfrag[b, lane, 4:8]
            Cm = A @ B                                                                 # exact s32
            assert np.abs(Cm).max() < 2 ** 24
            for lane in range(32):
                gid, tig = lane // 4, lane % 4
                if tig >= 2:
                    continue                                                          # columns 4..7: unused
                c0, c1, c2, c3 = Cm[gid, tig * 2], Cm[gid, tig * 2 + 1], Cm[gid + 8, tig * 2], Cm[gid + 8, tig * 2 + 1]
                w0, w1 = wj[tig * 2], wj[tig * 2 + 1]
                v0 = np.float32(np.float32(c0) * w0 + np.float32(c1) * w1)
                v8 = np.float32(np.float32(c2) * w0 + np.float32(c3) * w1)
                for j, (drow, v) in enumerate(((0, v0), (8, v8))):
                    row = R0 + gid + drow
                    dw = dplane[((row // 4) * (K // 32) + b) * 4 + (row % 4)]
                    acc[lane, j] = np.float32(acc[lane, j] + np.float32(dw * xs[b]) * v)
        for gid in range(8):
            for j, drow in enumerate((0, 8)):
                y[R0 + gid + drow] = acc[gid * 4 + 0, j] + acc[gid * 4 + 1, j]          # shfl_xor 1 over tig 0,1
    err = np.abs(y - ref).max() / np.abs(ref).max()
    return err


if __name__ == "__main__":
    for (M, K) in ((16, 32), (32, 128), (48, 256)):
        e = emulate(M, K)
        print("M=%d K=%d  max rel err %.2e" % (M, K, e))
        assert e < 2e-6
    print("ok")
