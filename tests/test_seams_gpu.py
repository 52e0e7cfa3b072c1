"""The primary drop-in seams of SURVEY 8(b), driven the way libvips drives them.

* vb200_reducev_gen / vb200_reduceh_gen: one call per sink tile (SMALLTILE 128 x 128 and FATSTRIP
  W x 16, non-zero left / top, shrink 1.7 -- not representable, so the per-rect coordinate stepping
  of reducev.cpp:548-611 / reduceh.cpp:254-326 shows), with the input region the reference's
  generate function would vips_region_prepare() on the embedded image.
* the scanline kernels vips_*_uchar_hwy (resample/presample.h:74-87, convolution/pconvolution.h:74-76):
  called from Python loops that restate vips_reducev_uchar_vector_gen (reducev.cpp:623-676),
  vips_reduceh_uchar_vector_gen (reduceh.cpp:340-392: virtual-origin p0, absolute X),
  vips_shrinkv_uchar_vector_gen (shrinkv.c:392-470), vips_shrinkh_uchar_vector_gen (shrinkh.c:290-353)
  and vips_convi_uchar_vector_gen (convi.c:306-362).

Expected pixels: tests/golden/seams_ref.npz, produced by the reference's own generate functions under
oracle/_ref with the same tiles (tests/golden/make_golden.py), and the oracle run with the same rects.
"""
import ctypes as C
import math
import os
import sys
import threading

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from cases import make_input, seam_cases  # noqa: E402

from oracle import pyconv, pyoracle as orc  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "seams_ref.npz"))
CASES = seam_cases()


def oracle_tiled(c):
    a = make_input(c)
    tw, th = c["tile"]
    if c["op"] == "reducev":
        return orc.reducev(a, c["f"], "lanczos3", 0.0, rect_h=th)
    return orc.reduceh(a, c["f"], "lanczos3", 0.0, rect_w=tw)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_tiled_reference(case):
    """CPU: the oracle, told the rect size, equals the reference pulled through those tiles."""
    assert np.array_equal(oracle_tiled(case), GOLD[case["name"]])


def test_tiling_matters_for_this_factor():
    """1.7 is not representable: SMALLTILE and FATSTRIP evaluations of the reference differ somewhere,
    so the seam tests below really do pin the per-rect stepping (if they did not differ, any tiling
    would pass)."""
    differ = [not np.array_equal(GOLD[CASES[i]["name"]], GOLD[CASES[i + 1]["name"]]) for i in range(0, len(CASES), 2)]
    assert any(differ)


# ------------------------------------------------------------------ helpers
def geometry(size, f):
    g = orc.reduce_geometry(size, f, "lanczos3", 0.0)
    assert g.int_shrink == 1
    return g.n_point, g.residual, g.offset, g.out_size


def embed_v(a, n_point):
    """vips_embed(in, 0, ceil(n_point / 2) - 1, w, h + n_point, extend copy), reducev.cpp:968-975"""
    top = int(math.ceil(n_point / 2.0) - 1)
    return np.ascontiguousarray(np.pad(a, ((top, n_point - top), (0, 0), (0, 0)), mode="edge"))


def embed_h(a, n_point):
    left = int(math.ceil(n_point / 2.0) - 1)
    return np.ascontiguousarray(np.pad(a, ((0, 0), (left, n_point - left), (0, 0)), mode="edge"))


def tiles(out_w, out_h, tw, th):
    tw = tw or out_w
    for y in range(0, out_h, th):
        for x in range(0, out_w, tw):
            yield x, y, min(tw, out_w - x), min(th, out_h - y)


def clip_rect(left, top, width, height, iw, ih):
    """vips_region_prepare clips the request to the image (region.c:1646-1690)"""
    r, b = min(left + width, iw), min(top + height, ih)
    left, top = max(left, 0), max(top, 0)
    return left, top, r - left, b - top


def region(vb, arr, left, top, full_w, full_h):
    from libvips_b200 import CImage, CRect, CRegion, FORMATS
    h, w, b = arr.shape
    im = CImage(full_w, full_h, b, FORMATS[arr.dtype], 22, vb.HOST, None, 0)
    return CRegion(im, CRect(left, top, w, h), arr.ctypes.data_as(C.c_void_p), arr.strides[0])


# ------------------------------------------------------------------ generate()-shaped seams
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reduce_gen_seams(vb, case):
    from libvips_b200 import CReduceParams
    a = make_input(case)
    H, W, B = a.shape
    vertical = case["op"] == "reducev"
    n_point, rs, off, out_size = geometry(H if vertical else W, case["f"])
    emb = embed_v(a, n_point) if vertical else embed_h(a, n_point)
    EH, EW = emb.shape[:2]
    out_w, out_h = (W, out_size) if vertical else (out_size, H)
    got = np.zeros((out_h, out_w, B), np.uint8)
    params = CReduceParams(n_point, 5, rs, off)
    fn = vb.lib().vb200_reducev_gen if vertical else vb.lib().vb200_reduceh_gen
    n_calls = 0
    for x, y, w, h in tiles(out_w, out_h, *case["tile"]):
        if vertical:  # reducev.cpp:539-544
            s = clip_rect(x, int(y * rs - off), w, int(h * rs + n_point), EW, EH)
        else:  # reduceh.cpp:240-245
            s = clip_rect(int(x * rs - off), y, int(w * rs + n_point), h, EW, EH)
        src = np.ascontiguousarray(emb[s[1]:s[1] + s[3], s[0]:s[0] + s[2]])
        dst = np.zeros((h, w, B), np.uint8)
        rin = region(vb, src, s[0], s[1], EW, EH)
        rout = region(vb, dst, x, y, out_w, out_h)
        vb._check(fn(C.byref(rout), C.byref(rin), C.byref(params)))
        got[y:y + h, x:x + w] = dst
        n_calls += 1
    assert n_calls > 1
    assert np.array_equal(got, GOLD[case["name"]]), "differs from the reference's own generate() on the same tiles"
    assert np.array_equal(got, oracle_tiled(case))


@pytest.mark.gpu
def test_gen_seams_refuse_short_regions(vb):
    from libvips_b200 import CReduceParams
    a = np.zeros((64, 64, 4), np.uint8)
    n_point, rs, off, out = geometry(64, 1.7)
    emb = embed_v(a, n_point)
    src = np.ascontiguousarray(emb[0:8])  # far too few rows for 16 output rows
    dst = np.zeros((16, 64, 4), np.uint8)
    rin = region(vb, src, 0, 0, 64, emb.shape[0])
    rout = region(vb, dst, 0, 0, 64, out)
    L = vb.lib()
    assert L.vb200_reducev_gen(C.byref(rout), C.byref(rin), C.byref(CReduceParams(n_point, 5, rs, off))) == -1
    assert b"does not cover" in L.vb200_error_buffer()
    L.vb200_error_clear()


# ------------------------------------------------------------------ scanline kernels, driven as the reference does
def addr(arr, y, x_bytes=0):
    return arr.ctypes.data + y * arr.strides[0] + x_bytes


def mask_tables(n_point, rs):
    _, s = orc.reduce_tables("lanczos3", n_point, rs)
    rows = [np.ascontiguousarray(s[i]) for i in range(65)]
    ptrs = (C.POINTER(C.c_short) * 65)(*[r.ctypes.data_as(C.POINTER(C.c_short)) for r in rows])
    return rows, ptrs


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c["op"] == "reduceh"], ids=lambda c: c["name"])
def test_vips_reduceh_uchar_hwy_as_the_vector_gen_calls_it(vb, case):
    """reduceh.cpp:340-392.  p0 is a VIRTUAL origin: tiles with left > 0 pass a pointer that lies
    before the region's buffer; only pixels from (int) X on may be read."""
    L = vb.lib()
    L.vips_reduceh_uchar_hwy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_double]
    L.vips_reduceh_uchar_hwy.restype = None
    a = make_input(case)
    H, W, B = a.shape
    n_point, rs, off, out_w = geometry(W, case["f"])
    emb = embed_h(a, n_point)
    EH, EW = emb.shape[:2]
    rows, cs = mask_tables(n_point, rs)
    got = np.zeros((H, out_w, B), np.uint8)
    for x, y, w, h in tiles(out_w, H, *case["tile"]):
        s = clip_rect(int(x * rs - off), y, int((w + 1) * rs + n_point), h, EW, EH)  # one extra pixel, :362
        # the region buffer: its own allocation, with a guard band before it that the kernel must not need
        ir = np.ascontiguousarray(emb[s[1]:s[1] + s[3], s[0]:s[0] + s[2]])
        q = np.zeros((h, w, B), np.uint8)
        for yy in range(h):
            X = (x + 0.5) * rs - 0.5 - off
            p0 = addr(ir, yy) - s[0] * B  # VIPS_REGION_ADDR(ir, ir->valid.left, y) - ir->valid.left * ps
            L.vips_reduceh_uchar_hwy(addr(q, yy), p0, n_point, w, B, cs, X, rs)
        got[y:y + h, x:x + w] = q
    assert L.vb200_error_buffer() == b""
    assert np.array_equal(got, GOLD[case["name"]])


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c["op"] == "reducev"], ids=lambda c: c["name"])
def test_vips_reducev_uchar_hwy_as_the_vector_gen_calls_it(vb, case):
    """reducev.cpp:623-676"""
    L = vb.lib()
    L.vips_reducev_uchar_hwy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.vips_reducev_uchar_hwy.restype = None
    a = make_input(case)
    H, W, B = a.shape
    n_point, rs, off, out_h = geometry(H, case["f"])
    emb = embed_v(a, n_point)
    EH, EW = emb.shape[:2]
    rows, _ = mask_tables(n_point, rs)
    got = np.zeros((out_h, W, B), np.uint8)
    for x, y, w, h in tiles(W, out_h, *case["tile"]):
        s = clip_rect(x, int(y * rs - off), w, int(h * rs + n_point), EW, EH)
        ir = np.ascontiguousarray(emb[s[1]:s[1] + s[3], s[0]:s[0] + s[2]])
        q = np.zeros((h, w, B), np.uint8)
        Y = (y + 0.5) * rs - 0.5 - off
        for yy in range(h):
            py = int(Y)
            sy = int(Y * 64 * 2)
            ty = ((sy & 127) + 1) >> 1
            L.vips_reducev_uchar_hwy(addr(q, yy), addr(ir, py - s[1], (x - s[0]) * B), n_point, w * B, ir.strides[0],
                                     rows[ty].ctypes.data)
            Y += rs
        got[y:y + h, x:x + w] = q
    assert L.vb200_error_buffer() == b""
    assert np.array_equal(got, GOLD[case["name"]])


@pytest.mark.gpu
def test_vips_shrink_uchar_hwy_as_the_vector_gens_call_them(vb):
    """shrinkv.c:392-470 (add_line into a per-chunk sum buffer, then write_line) and shrinkh.c:290-353."""
    L = vb.lib()
    L.vips_shrinkv_add_line_uchar_hwy.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.vips_shrinkv_write_line_uchar_hwy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.vips_shrinkh_uchar_hwy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    for f in (L.vips_shrinkv_add_line_uchar_hwy, L.vips_shrinkv_write_line_uchar_hwy, L.vips_shrinkh_uchar_hwy):
        f.restype = None
    rng = np.random.default_rng(31)
    a = rng.integers(0, 256, (96, 150, 4), dtype=np.uint8)
    H, W, B = a.shape
    dy = 16  # vips__fatstrip_height

    # --- shrinkv, factor 3, SMALLTILE-like rects with left / top > 0
    vs = 3
    want = orc.shrinkv(a, vs)
    oh = want.shape[0]
    got = np.zeros_like(want)
    for x, y, w, h in tiles(W, oh, 64, 24):
        ne = w * B
        for c0 in range(0, h, dy):
            ch = min(dy, h - c0)
            sums = np.zeros((dy, ne), np.uint32)  # memset(seq->sum, 0, ...)
            start, end = (y + c0) * vs, (y + c0 + ch) * vs
            for y1 in range(start, end):
                row = np.ascontiguousarray(a[y1, x:x + w])
                L.vips_shrinkv_add_line_uchar_hwy(row.ctypes.data, ne, addr(sums, (y1 - start) // vs))
            for y1 in range(ch):
                q = np.zeros(ne, np.uint8)
                L.vips_shrinkv_write_line_uchar_hwy(q.ctypes.data, ne, vs, addr(sums, y1))
                got[y + c0 + y1, x:x + w] = q.reshape(w, B)
    assert np.array_equal(got, want)

    # --- shrinkh, factor 5 (150 / 5 = 30 columns), rects with left > 0
    hs = 5
    want = orc.shrinkh(a, hs)
    ow = want.shape[1]
    got = np.zeros_like(want)
    for x, y, w, h in tiles(ow, H, 16, 16):
        for yy in range(h):
            # s.left = r->left * hshrink; one extra pixel is requested but the kernel may not need it
            row = np.ascontiguousarray(a[y + yy, x * hs:min(W, (x + w + 1) * hs)])
            q = np.zeros((w, B), np.uint8)
            L.vips_shrinkh_uchar_hwy(q.ctypes.data, row.ctypes.data, w, hs, B)
            got[y + yy, x:x + w] = q
    assert L.vb200_error_buffer() == b""
    assert np.array_equal(got, want)


class GObjectHead(C.Structure):
    _fields_ = [("g_class", C.c_void_p), ("ref_count", C.c_uint), ("qdata", C.c_void_p)]


class VipsObjectHead(C.Structure):
    _fields_ = [("parent_instance", GObjectHead), ("constructed", C.c_int), ("static_object", C.c_int),
                ("argument_table", C.c_void_p), ("nickname", C.c_char_p), ("description", C.c_char_p),
                ("preclose", C.c_int), ("close", C.c_int), ("postclose", C.c_int), ("local_memory", C.c_size_t)]


class VipsImageHead(C.Structure):
    _fields_ = [("parent_instance", VipsObjectHead), ("Xsize", C.c_int), ("Ysize", C.c_int), ("Bands", C.c_int),
                ("BandFmt", C.c_int), ("Coding", C.c_int), ("Type", C.c_int)]


class VipsRect(C.Structure):
    _fields_ = [("left", C.c_int), ("top", C.c_int), ("width", C.c_int), ("height", C.c_int)]


class VipsRegionHead(C.Structure):
    _fields_ = [("parent_object", VipsObjectHead), ("im", C.POINTER(VipsImageHead)), ("valid", VipsRect),
                ("type", C.c_int), ("data", C.c_void_p), ("bpl", C.c_int), ("seq", C.c_void_p)]


def test_vips_abi_mirror_layout():
    """CPU: the LP64 offsets include/vb200_vips_abi.h asserts (GObject 24 B, VipsObject 80 B)."""
    assert C.sizeof(GObjectHead) == 24 and C.sizeof(VipsObjectHead) == 80
    assert VipsImageHead.Bands.offset == 88
    assert (VipsRegionHead.im.offset, VipsRegionHead.valid.offset, VipsRegionHead.data.offset,
            VipsRegionHead.bpl.offset) == (80, 88, 112, 120)


@pytest.mark.gpu
def test_vips_convi_uchar_hwy_as_the_vector_gen_calls_it(vb):
    """convi.c:306-362 + pconvolution.h:74-76: VipsRegion pointers, element offsets computed from the
    input region's own bpl, the 8-bit-mantissa mask of vips_convi_intize."""
    L = vb.lib()
    L.vips_convi_uchar_hwy.restype = None
    L.vips_convi_uchar_hwy.argtypes = [C.POINTER(VipsRegionHead), C.POINTER(VipsRegionHead), C.POINTER(VipsRect), C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(77)
    a = rng.integers(0, 256, (90, 110, 3), dtype=np.uint8)
    H, W, B = a.shape
    mask = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1], [0, 3, 0], [2, 0, 2]], np.float64)
    scale, offset = 23.0, 2.0
    mh, mw = mask.shape
    iz = pyconv.convi_intize8(mask, scale)
    assert iz is not None
    mant, pos, exp = iz
    mant = np.ascontiguousarray(mant, np.int16)
    want = pyconv.conv(a, mask, scale=scale, offset=offset, precision="integer", vector=True)
    emb = np.ascontiguousarray(np.pad(a, ((mh // 2, mh - 1 - mh // 2), (mw // 2, mw - 1 - mw // 2), (0, 0)), mode="edge"))
    in_im = VipsImageHead(Xsize=emb.shape[1], Ysize=emb.shape[0], Bands=B, BandFmt=0, Coding=0, Type=22)
    out_im = VipsImageHead(Xsize=W, Ysize=H, Bands=B, BandFmt=0, Coding=0, Type=22)
    got = np.zeros_like(a)
    for x, y, w, h in tiles(W, H, 48, 40):
        s = (x, y, w + mw - 1, h + mh - 1)
        # a region buffer wider than the rect (bpl > line), as a VipsBuffer may be
        buf = np.zeros((s[3], s[2] + 5, B), np.uint8)
        buf[:, :s[2]] = emb[s[1]:s[1] + s[3], s[0]:s[0] + s[2]]
        q = np.zeros((h, w, B), np.uint8)
        ir = VipsRegionHead(im=C.pointer(in_im), valid=VipsRect(*s), type=1, data=buf.ctypes.data, bpl=buf.strides[0])
        orr = VipsRegionHead(im=C.pointer(out_im), valid=VipsRect(x, y, w, h), type=1, data=q.ctypes.data, bpl=q.strides[0])
        # seq->offsets[i], convi.c:336-347
        offs = np.array([(int(p) // mw) * buf.strides[0] + (int(p) % mw) * B for p in pos], np.int32)
        r = VipsRect(x, y, w, h)
        L.vips_convi_uchar_hwy(C.byref(orr), C.byref(ir), C.byref(r), w * B, len(pos), int(round(offset)), offs.ctypes.data,
                               mant.ctypes.data, exp)
        got[y:y + h, x:x + w] = q
    assert L.vb200_error_buffer() == b""
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ concurrent callers
@pytest.mark.gpu
def test_four_concurrent_caller_threads(vb):
    """generate() callbacks arrive concurrently from worker threads (iofuncs/region.c:1600-1626,
    doc/using-threads.md): four host threads hammer different seams of the one library at once;
    every result must equal the single-threaded answer, and errors must stay per thread."""
    from libvips_b200 import CReduceParams
    rng = np.random.default_rng(404)
    a = rng.integers(0, 256, (640, 512, 4), dtype=np.uint8)
    rgb = np.ascontiguousarray(a[:, :, :3])
    want_thumb = orc.thumbnail_image(a, 96)
    want_lab = orc.colourspace(rgb, "lab", "srgb")
    want_blur = pyconv.gaussblur(rgb, 1.2, 0.2, "integer")
    want_sharp = pyconv.sharpen(rgb, "srgb")
    n_point, rs, off, out_h = geometry(640, 1.7)
    emb = embed_v(a, n_point)
    want_rv = orc.reducev(a, 1.7, "lanczos3", 0.0, rect_h=16)
    plan = vb.ThumbnailPlan(512, 640, 4, 96)
    frames = np.stack([a, a[::-1].copy(), np.roll(a, 5, 1)])
    want_batch = np.stack([orc.thumbnail_image(f, 96) for f in frames])
    errors = []

    def reducev_tiles():
        got = np.zeros_like(want_rv)
        params = CReduceParams(n_point, 5, rs, off)
        for x, y, w, h in tiles(512, out_h, 0, 16):
            s = clip_rect(x, int(y * rs - off), w, int(h * rs + n_point), emb.shape[1], emb.shape[0])
            src = np.ascontiguousarray(emb[s[1]:s[1] + s[3]])
            dst = np.zeros((h, w, 4), np.uint8)
            rin = region(vb, src, s[0], s[1], emb.shape[1], emb.shape[0])
            rout = region(vb, dst, x, y, 512, out_h)
            vb._check(vb.lib().vb200_reducev_gen(C.byref(rout), C.byref(rin), C.byref(params)))
            got[y:y + h] = dst
        assert np.array_equal(got, want_rv)

    def worker(k):
        try:
            for it in range(6):
                job = (k + it) % 6
                if job == 0:
                    assert np.array_equal(vb.Image(a).thumbnail_image(96).numpy(), want_thumb)
                elif job == 1:
                    assert np.array_equal(vb.Image(rgb).colourspace("lab").numpy(), want_lab)
                elif job == 2:
                    assert np.array_equal(vb.Image(rgb).gaussblur(1.2).numpy(), want_blur)
                elif job == 3:
                    assert np.array_equal(vb.Image(rgb).sharpen().numpy(), want_sharp)
                elif job == 4:
                    reducev_tiles()
                else:
                    assert np.array_equal(plan.run_host(frames), want_batch)  # one shared plan: its pump serialises
                # an error raised on this thread must not leak into another thread's buffer
                with pytest.raises(vb.Error, match="reduce factor should be >= 1.0"):
                    vb.Image(rgb).reduce(0.5, 0.5)
                assert vb.lib().vb200_error_buffer() == b""
        except BaseException as e:  # noqa: BLE001 -- reported by the main thread
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


# ------------------------------------------------------------------ vips_reduce with the caller's factors
@pytest.mark.gpu
@pytest.mark.parametrize("hs,vs", [(49.0, 49.0), (2.37, 2.37), (93.0, 1.0), (1.0, 3.3), (4.75, 2.5)])
def test_reduce_uses_the_callers_factors(vb, hs, vs):
    """reduce.c:106-114: reducev(vshrink) then reduceh(hshrink) -- not resize(1 / h, 1 / v), whose second
    inversion changes ~13% of factors (1 / (1 / 49) != 49)."""
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    want = orc.reduceh(orc.reducev(a, vs, "lanczos3", 0.0, rect_h=16), hs, "lanczos3", 0.0, rect_w=0)
    got = vb.Image(a).reduce(hs, vs, gap=0.0).numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_reduce_rejects_enlarging_factors(vb):
    a = np.zeros((32, 32, 3), np.uint8)
    with pytest.raises(vb.Error, match="reduce factor should be >= 1.0"):
        vb.Image(a).reduce(0.5, 2.0)
    with pytest.raises(vb.Error, match="reduce factor should be >= 1.0"):
        vb.Image(a).reduce(2.0, 0.5)
