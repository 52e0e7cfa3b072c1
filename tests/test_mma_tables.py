"""Host logic of the tensor-pipe reducev (thumbnail_fused_mma.cuh), on the CPU.

The plan places the reducev coefficients into mma.m16n8k32 B fragments by ring slot
(build_mma_tables).  This test replays what the kernel does with those tables -- quads of 4
box-shrunk rows in a ring of 8, D = A x B with the coefficients split hi * 256 + lo,
(sum + 2048) >> 12, clip -- in numpy, and requires the oracle's shrinkv + reducev bytes."""
import ctypes as C

import numpy as np
import pytest

import libvips_b200 as vb
from oracle import pyoracle as orc


def tables(in_size, shrink, rect=16):
    L = C.CDLL(vb.library_path())
    cap = 4096
    ints = [C.c_int() for _ in range(5)]
    rows = C.c_int(0)
    first = (C.c_int * cap)()
    phase = (C.c_int * cap)()
    mask = (C.c_short * (65 * 64))()
    vchunk = (C.c_int * (2 * cap // 8 + 2))()
    bfrag = (C.c_uint * (128 * (cap // 8 + 1)))()
    L.vb200_debug_mma_tables.argtypes = [C.c_int, C.c_double, C.c_int] + [C.POINTER(C.c_int)] * 7 + \
        [C.POINTER(C.c_short), C.POINTER(C.c_int), C.POINTER(C.c_uint), C.c_int, C.POINTER(C.c_int)]
    rc = L.vb200_debug_mma_tables(in_size, shrink, rect, *[C.byref(i) for i in ints], first, phase, mask, vchunk, bfrag, cap, C.byref(rows))
    vs, hs, oh, npnt, embed = [i.value for i in ints]
    R = max(rows.value, 1)
    nch = (oh + R - 1) // R
    return rc, dict(VS=vs, Hs=hs, OH=oh, n_point=npnt, embed=embed, rows=R, first=np.array(first[:oh]), phase=np.array(phase[:oh]),
                    vchunk=np.array(vchunk[:2 * nch]).reshape(-1, 2),
                    bfrag=np.array(bfrag[:128 * nch], dtype=np.uint32).reshape(-1, 32, 4))


@pytest.mark.parametrize("in_size,shrink", [(4096, 8.0), (2048, 8.0), (1024, 4.0), (1600, 8.0), (1000, 4.0), (4096, 8.7),
                                            (2000, 4.76), (3000, 5.9), (4096, 9.9), (4096, 16.0), (2160, 17.3)])
def test_mma_tables_reproduce_reducev(in_size, shrink):
    rc, t = tables(in_size, shrink)
    assert rc == 0, "some chunking of 4..8 rows must fit the 8-quad ring for shrinks of 4..10"
    assert t["VS"] in (2, 4, 8)
    rng = np.random.default_rng(in_size)
    col = rng.integers(0, 256, (in_size, 3, 1), dtype=np.uint8)       # a 3-pixel-wide, 1-band image
    box = orc.shrinkv(col, t["VS"], ceil=True)[:t["Hs"]]              # what the V warps average
    want = orc.reducev(col if t["VS"] == 1 else box, shrink / t["VS"], "lanczos3", gap=0.0, rect_h=16)
    # reducev of the box-shrunk image at the residual factor is exactly the second half of vips_reduce(gap 2)
    assert want.shape[0] == t["OH"]
    ring = np.zeros((8, 4, 3), np.int64)                                # [slot][row in quad][column]
    produced = t["vchunk"][0, 0] - 1
    got = np.zeros_like(want)
    for c, (q0, q1) in enumerate(t["vchunk"]):
        assert q1 - q0 < 8 and q0 <= produced + 1
        for q in range(produced + 1, q1 + 1):                            # the V warps, in order, no gaps
            for i in range(4):
                src = min(max(4 * q + i - t["embed"], 0), t["Hs"] - 1)   # edge rows replicate
                ring[q & 7, i] = box[src, :, 0]
        produced = max(produced, q1)
        for lane in range(32):                                           # the MMA: lane (g, tig) holds B[k][n = g]
            tig, g = lane & 3, lane >> 2
            y = c * t["rows"] + g
            if g >= t["rows"] or y >= t["OH"]:
                assert not t["bfrag"][c, lane].any() or y < t["OH"]
                continue
            w = t["bfrag"][c, lane]
            acc = np.full(3, 2048, np.int64)
            for half in range(2):
                slot = tig + 4 * half
                for i in range(4):
                    hi = (int(w[half]) >> (8 * i)) & 0xff
                    hi -= 256 if hi >= 128 else 0                        # s8
                    lo = (int(w[2 + half]) >> (8 * i)) & 0xff            # u8
                    acc += (hi * 256 + lo) * ring[slot, i]
            # the four lanes with the same g hold different k ranges: the MMA sums over all four tig
            if tig == 0:
                total = np.zeros(3, np.int64)
            total = total + acc - 2048
            if tig == 3:
                got[y, :, 0] = np.clip((total + 2048) >> 12, 0, 255)
    assert np.array_equal(got, want)


def test_mma_tables_rows_per_chunk():
    """exact factors keep 8 rows per chunk; a residual shrink of 2.4 (17 taps at stride 2.4) needs fewer"""
    assert tables(4096, 8.0)[1]["rows"] == 8
    rc, t = tables(2000, 4.76)
    assert rc == 0 and 4 <= t["rows"] < 8


def replay(in_size, shrink):
    """the kernel's use of the tables, against the oracle's whole vips_reducev(gap 2) -- box shrink by any VS (2 .. 8, the
    non-powers of two included), ceil'd heights, the output height from the ORIGINAL size: None when the plan falls back"""
    rc, t = tables(in_size, shrink)
    if rc != 0:
        return None
    rng = np.random.default_rng(in_size)
    col = rng.integers(0, 256, (in_size, 3, 1), dtype=np.uint8)
    box = orc.shrinkv(col, t["VS"], ceil=True)[:t["Hs"]] if t["VS"] > 1 else col
    want = orc.reducev(col, shrink, "lanczos3", gap=2.0, rect_h=16)
    assert want.shape[0] == t["OH"], (in_size, shrink)
    ring = np.zeros((8, 4, 3), np.int64)
    produced = t["vchunk"][0, 0] - 1
    got = np.zeros_like(want)
    for c, (q0, q1) in enumerate(t["vchunk"]):
        assert q1 - q0 < 8 and q0 <= produced + 1, (in_size, shrink, c)
        for q in range(produced + 1, q1 + 1):
            for i in range(4):
                ring[q & 7, i] = box[min(max(4 * q + i - t["embed"], 0), t["Hs"] - 1), :, 0]
        produced = max(produced, q1)
        for g in range(min(t["rows"], 8)):
            y = c * t["rows"] + g
            if y >= t["OH"]:
                continue
            total = np.zeros(3, np.int64)
            for tig in range(4):
                w = t["bfrag"][c, g * 4 + tig]
                for half in range(2):
                    for i in range(4):
                        hi = (int(w[half]) >> (8 * i)) & 0xff
                        hi -= 256 if hi >= 128 else 0
                        lo = (int(w[2 + half]) >> (8 * i)) & 0xff
                        total += (hi * 256 + lo) * ring[tig + 4 * half, i]
            got[y, :, 0] = np.clip((total + 2048) >> 12, 0, 255)
    assert np.array_equal(got, want), (in_size, shrink)
    return t


def test_mma_tables_random_geometries():
    """200 random (size, shrink) pairs, shrinks 4 .. 18 incl. the boxes 3 / 5 / 6 / 7 and sizes that divide by nothing: every
    plan the tensor-pipe kernel accepts reproduces the oracle (a longer run of the same loop: 6 153 geometries, 60 fallbacks,
    no mismatch)"""
    rng = np.random.default_rng(2024)
    seen, fallback = set(), 0
    for _ in range(200):
        in_size = int(rng.integers(64, 5000))
        shrink = float(rng.choice([rng.random() * 14 + 4, rng.integers(4, 19), round(rng.random() * 14 + 4, 2)]))
        if in_size / shrink < 2:
            continue
        t = replay(in_size, shrink)
        if t is None:
            fallback += 1
        else:
            seen.add(t["VS"])
    assert {2, 3, 4, 5, 6, 7, 8} <= seen and fallback < 20, (seen, fallback)
