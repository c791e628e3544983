#!/usr/bin/env python
"""CPU emulation of gemv_phase_q8 / q8_digits_phase / q8_tile_partial of the Q8 megakernel experiment
(kernels_mega.cu, LB_Q8_MEGA=1): shared-memory B-fragment word layout, the K split over the CTA's 16 warps (ragged
block ranges), 16-row tiles, the combine over warps.  Checks the result against d*q.x in f64."""
import numpy as np

from q8_mma_layout_emulation import digits_of_block, quantize_interleaved

MG_WARPS = 16


def digits_phase(vec):
    """-> bfrag words [NB][32] (uint32) and xsc [NB], written exactly as q8_digits_phase does."""
    K = vec.size
    NB = K // 32
    bfrag = np.zeros((NB, 32), np.uint32)
    xsc = np.zeros(NB, np.float32)
    for b in range(NB):
        xsc[b], digs = digits_of_block(vec[32 * b:32 * b + 32])
        for lane in range(32):
            tig, reg = (lane & 15) >> 2, lane >> 4
            for j in range(4):
                byte = np.uint32(np.uint8(digs[j, lane])) << np.uint32(8 * (lane & 3))
                bfrag[b, (j * 4 + tig) * 2 + reg] |= byte          # the 4 lanes of a quad OR their bytes into one word
    return bfrag, xsc


def tile_partial(qwords, dplane, R0, K, b_begin, b_end, bfrag, xsc):
    """one warp: per-lane (acc_lo, acc_hi) after the reduction over the quad, as q8_tile_partial returns them"""
    NB, K4 = K // 32, K // 4
    acc = np.zeros((32, 2), np.float32)
    for lane in range(32):
        gid, tig = lane >> 2, lane & 3
        r_lo, r_hi = R0 + gid, R0 + gid + 8
        qa = ((r_lo >> 2) * K4 + tig) * 4 + (r_lo & 3)
        qb = ((r_hi >> 2) * K4 + tig) * 4 + (r_hi & 3)
        da = (r_lo >> 2) * NB * 4 + (r_lo & 3)
        db = (r_hi >> 2) * NB * 4 + (r_hi & 3)
        w0 = np.float32(2.0 ** -7 if tig == 0 else 2.0 ** -21)
        w1 = np.float32(w0 * np.float32(2.0 ** -7))
        for b in range(b_begin, b_end):
            w = b * 32
            a = [qwords[qa + w], qwords[qb + w], qwords[qa + w + 16], qwords[qb + w + 16]]
            bf = (bfrag[b, (gid * 4 + tig) * 2], bfrag[b, (gid * 4 + tig) * 2 + 1]) if gid < 4 else (np.uint32(0), np.uint32(0))
            acc[lane, 0] += 0  # placeholder: the MMA needs all lanes; done below per warp
    # the MMA is a warp-wide operation: gather fragments of all lanes per block, multiply, scatter C
    for b in range(b_begin, b_end):
        A = np.zeros((16, 32), np.int64); B = np.zeros((32, 8), np.int64)
        for lane in range(32):
            gid, tig = lane >> 2, lane & 3
            r_lo, r_hi = R0 + gid, R0 + gid + 8
            qa = ((r_lo >> 2) * K4 + tig) * 4 + (r_lo & 3)
            qb = ((r_hi >> 2) * K4 + tig) * 4 + (r_hi & 3)
            w = b * 32
            for reg, (word, row, k0) in enumerate(((qa + w, gid, tig * 4), (qb + w, gid + 8, tig * 4),
                                                   (qa + w + 16, gid, 16 + tig * 4), (qb + w + 16, gid + 8, 16 + tig * 4))):
                A[row, k0:k0 + 4] = qwords[word:word + 1].view(np.int8)
            if gid < 4:
                b0 = bfrag[b, (gid * 4 + tig) * 2:(gid * 4 + tig) * 2 + 1].view(np.int8)
                b1 = bfrag[b, (gid * 4 + tig) * 2 + 1:(gid * 4 + tig) * 2 + 2].view(np.int8)
                B[tig * 4:tig * 4 + 4, gid] = b0
                B[16 + tig * 4:16 + tig * 4 + 4, gid] = b1
        Cm = A @ B
        for lane in range(32):
            gid, tig = lane >> 2, lane & 3
            r_lo, r_hi = R0 + gid, R0 + gid + 8
            w0 = np.float32(2.0 ** -7 if tig == 0 else 2.0 ** -21); w1 = np.float32(w0 * np.float32(2.0 ** -7))
            c0, c1, c2, c3 = Cm[gid, tig * 2], Cm[gid, tig * 2 + 1], Cm[gid + 8, tig * 2], Cm[gid + 8, tig * 2 + 1]
            v_lo = np.float32(np.float32(c0) * w0 + np.float32(c1) * w1)
            v_hi = np.float32(np.float32(c2) * w0 + np.float32(c3) * w1)
            s_lo = dplane[(r_lo >> 2) * NB * 4 + (r_lo & 3) + b * 4]
            s_hi = dplane[(r_hi >> 2) * NB * 4 + (r_hi & 3) + b * 4]
            acc[lane, 0] = np.float32(acc[lane, 0] + np.float32(s_lo * xsc[b]) * v_lo)
            acc[lane, 1] = np.float32(acc[lane, 1] + np.float32(s_hi * xsc[b]) * v_hi)
    out = np.zeros((32, 2), np.float32)
    for lane in range(32):                       # shfl_xor 1, 2: every lane of the quad gets the quad's sum
        base = lane & ~3
        out[lane] = acc[base] + acc[base + 1] + acc[base + 2] + acc[base + 3]
    return out


def gemv_phase(W, x):
    M, K = W.shape
    q, d, qplane, dplane = quantize_interleaved(W)
    ref = (d.repeat(32, axis=1).astype(np.float64) * q) @ x.astype(np.float64)
    qwords = qplane.view(np.int32)
    bfrag, xsc = digits_phase(x)
    NB = K // 32
    per = (NB + MG_WARPS - 1) // MG_WARPS
    y = np.zeros(M, np.float32)
    for tile in range(M // 16):
        part = np.zeros((16, MG_WARPS), np.float32)
        for warp in range(MG_WARPS):
            b_begin = min(warp * per, NB); b_end = min(b_begin + per, NB)
            res = tile_partial(qwords, dplane, tile * 16, K, b_begin, b_end, bfrag, xsc)
            for gid in range(8):                  # tig == 0 lanes write the partials
                part[gid, warp] = res[gid * 4, 0]
                part[gid + 8, warp] = res[gid * 4, 1]
        for r in range(16):
            s1 = np.float32(0)
            for wv in range(MG_WARPS):
                s1 = np.float32(s1 + part[r, wv])
            y[tile * 16 + r] = s1
    return np.abs(y - ref).max() / np.abs(ref).max()


if __name__ == "__main__":
    rs = np.random.RandomState(1)
    for (M, K) in ((16, 64), (32, 32 * 37), (16, 32 * 16)):
        W = (rs.randn(M, K) / 8).astype(np.float32)
        x = (rs.randn(K) * np.exp(rs.randn(K))).astype(np.float32)
        e = gemv_phase(W, x)
        print("M=%d K=%d (NB=%d)  max rel err %.2e" % (M, K, K // 32, e))
        assert e < 2e-6
    print("ok")
