"""Host-side checks of the work split and the data layouts of the ring megakernels (round 2), evaluated through
`lb_layout_query` by the SAME host/device functions the kernels use (no GPU needed):

* FP32 ring (kernels_ring.cu): a row of K floats is streamed as chunks of <= 4096 floats, every chunk a multiple of 64
  floats so that each of the 16 consumer warps owns whole float4s;
* Q8 ring (kernels_ring_q8.cu): the rows of a matrix are dealt to 148 work slots exactly once; the decode plane stores
  every tile (<= 16 rows) as compact, 16-byte aligned records that tile the plane without gaps or overlap;
* pods ring (kernels_ring_pods.cu): rows dealt exactly once; a slot that uses the chunk-bounded tensor map owns exactly
  chunk c of ceil(M / 148) rows.
"""
import ctypes as C

import pytest

import llama_go_b200  # noqa: F401
from llama_go_b200 import _capi

GRID = 148
SHAPES_7B = [(12288, 4096), (4096, 4096), (11008, 4096), (4096, 11008), (32000, 4096)]
SHAPES_13B = [(15360, 5120), (5120, 5120), (13824, 5120), (5120, 13824)]
SHAPES_65B = [(24576, 8192), (8192, 8192), (22016, 8192), (8192, 22016)]
SHAPES_SMALL = [(768, 256), (256, 256), (704, 256), (1024, 256), (2304, 768), (2048, 768), (768, 2048)]


def query(kind, a, b=0, c=0):
    out = (C.c_uint32 * 4)()
    rc = _capi.lib().lb_layout_query(kind, a, b, c, out)
    return rc, list(out)


@pytest.mark.parametrize("K", [256, 704, 768, 2048, 4096, 5120, 8192, 11008, 13824, 17920, 22016])
def test_fp32_ring_chunks(K):
    rc, (nch, ch, _, _) = query(0, K)
    assert rc == 0 and nch == -(-K // 4096)
    assert ch % 64 == 0 and 0 < ch <= 4096
    lens = [min(ch, K - c * ch) for c in range(nch)]
    assert all(l > 0 and l % 64 == 0 for l in lens) and sum(lens) == K     # every warp's 1/16 slice is whole float4s
    assert all((l // 16) <= 256 for l in lens)                              # two float4 per lane cover a slice


@pytest.mark.parametrize("M,K", SHAPES_7B + SHAPES_13B + SHAPES_65B + SHAPES_SMALL)
def test_q8_rows_are_dealt_exactly_once_and_balanced(M, K):
    if query(1, M, K, 0)[0] != 0:
        pytest.skip("shape not taken by the Q8 ring")
    seen, sizes = [], []
    for c in range(GRID):
        rc, (r0, r1, _, _) = query(1, M, K, c)
        assert rc == 0 and r0 <= r1 <= M
        seen += list(range(r0, r1))
        sizes.append(r1 - r0)
    assert seen == list(range(M))
    if M >= 16 * GRID:                               # row-balanced chunks: nobody has more than ceil(M / 148) rows
        assert max(sizes) == -(-M // GRID)


@pytest.mark.parametrize("M,K", SHAPES_7B + SHAPES_13B + SHAPES_SMALL)
def test_q8_decode_plane_records_tile_the_plane(M, K):
    if query(2, M, K, 0)[0] != 0:
        pytest.skip("shape not taken by the Q8 ring")
    nblk = K // 32
    tiles = {}
    for row in range(M):
        rc, (g0, rt, lo, hi) = query(2, M, K, row)
        assert rc == 0 and g0 <= row < g0 + rt and 1 <= rt <= 16
        tiles[g0] = (rt, lo | (hi << 32))
    # tiles partition the rows; a tile never straddles two work slots
    starts = sorted(tiles)
    assert starts[0] == 0 and all(starts[i] + tiles[starts[i]][0] == starts[i + 1] for i in range(len(starts) - 1))
    assert starts[-1] + tiles[starts[-1]][0] == M
    bounds = set()
    for c in range(GRID):
        _, (r0, r1, _, _) = query(1, M, K, c)
        bounds.update((r0, r1))
    for g0, (rt, _) in tiles.items():
        assert not any(g0 < b < g0 + rt for b in bounds)
    # records: (tile, 1024-column segment) -> [nb][rt][32] int8 | [nb][rt] f32, contiguous, 16-byte aligned, no gaps
    end = 0
    for g0 in starts:
        rt, off = tiles[g0]
        assert off == end == g0 * nblk * 36
        for seg in range(-(-K // 1024)):
            nb = min(32, nblk - seg * 32)
            rec = off + seg * 32 * rt * 36
            size = nb * rt * 36
            assert rec % 16 == 0 and size % 16 == 0 and size <= 18432      # one bulk copy, one ring slot
        end = off + nblk * rt * 36
    assert end == M * nblk * 36                                           # = q8_tile_major_bytes


@pytest.mark.parametrize("M", [256, 704, 1024, 4096, 5120, 8192, 11008, 12288, 13824, 22016, 32000])
def test_pods_rows_and_chunks(M):
    seen = []
    R = -(-M // GRID)
    for c in range(GRID):
        rc, (r0, r1, chunk, _) = query(3, M, 0, c)
        assert rc == 0 and r0 <= r1 <= M
        seen += list(range(r0, r1))
        if chunk != 0xFFFFFFFF:                       # chunk-bounded tensor map: exactly chunk c, entirely inside the matrix
            assert M >= 16 * GRID and chunk == c and r0 == c * R and r1 == r0 + R
        if M >= 16 * GRID:
            assert r1 - r0 <= R
        else:                                         # small matrices: whole 16-row tiles through the matrix-bounded map
            assert r0 % 16 == 0 and (r1 % 16 == 0 or r1 == M)
    assert seen == list(range(M))
