"""The generate()-shaped seams of SURVEY 8(b): vb200_{shrinkv,shrinkh,conv,colour}_gen fill
one output rect from an input region, the way a VipsGenerateFn is called.  Each is checked
against the whole-image oracle result cropped to the same rect."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyconv, pyoracle as orc

pytestmark = pytest.mark.gpu


def region(vb, arr, left, top, interp=22):
    """A CRegion over `arr` (H x W x B) whose pixel (0, 0) sits at image position (left, top)."""
    from libvips_b200 import CImage, CRect, CRegion, FORMATS
    h, w, b = arr.shape
    im = CImage(w + left, h + top, b, FORMATS[arr.dtype], interp, vb.HOST, None, 0)
    return CRegion(im, CRect(left, top, w, h), arr.ctypes.data_as(C.c_void_p), arr.strides[0])


def test_shrink_gens(vb):
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (96, 128, 4), dtype=np.uint8)
    L = vb.lib()
    # rows 8..23 of shrinkv(a, 3) need input rows 24..71
    want = orc.shrinkv(a, 3)[8:24]
    out = np.zeros_like(want)
    src = np.ascontiguousarray(a[24:72])  # the CRegion holds a raw pointer: keep the array alive
    rin = region(vb, src, 0, 24)
    rout = region(vb, out, 0, 8)
    vb._check(L.vb200_shrinkv_gen(C.byref(rout), C.byref(rin), 3))
    assert np.array_equal(out, want)
    # columns 5..24 of shrinkh(a, 4) need input columns 20..99
    want = np.ascontiguousarray(orc.shrinkh(a, 4)[:, 5:25])
    out = np.zeros_like(want)
    src = np.ascontiguousarray(a[:, 20:100])
    rin = region(vb, src, 20, 0)
    rout = region(vb, out, 5, 0)
    vb._check(L.vb200_shrinkh_gen(C.byref(rout), C.byref(rin), 4))
    assert np.array_equal(out, want)


@pytest.mark.parametrize("precision", ["float", "integer"])
def test_conv_gen(vb, precision):
    from libvips_b200 import CMask, PRECISIONS
    rng = np.random.default_rng(6)
    a = rng.integers(0, 256, (80, 90, 3), dtype=np.uint8)
    mask = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1], [0, 1, 0], [1, 0, 1]], np.float64)  # 3 wide, 5 high
    mh, mw = mask.shape
    want = pyconv.conv(a, mask, scale=19.0, offset=1.0, precision=precision)[20:36, 30:62]
    # embedded image = a edge-extended by (mw / 2, mh / 2); output (x, y) reads embedded (x.., y..)
    emb = np.pad(a, ((mh // 2, mh - 1 - mh // 2), (mw // 2, mw - 1 - mw // 2), (0, 0)), mode="edge")
    rin_arr = np.ascontiguousarray(emb[20:36 + mh - 1, 30:62 + mw - 1])
    out = np.zeros_like(want)
    cm = CMask(mw, mh, mask.ctypes.data_as(C.POINTER(C.c_double)), 19.0, 1.0)
    rin = region(vb, rin_arr, 30, 20)
    rout = region(vb, out, 30, 20)
    vb._check(vb.lib().vb200_conv_gen(C.byref(rout), C.byref(rin), C.byref(cm), PRECISIONS[precision]))
    assert np.array_equal(out, want)


def test_colour_gen(vb):
    rng = np.random.default_rng(7)
    a = rng.integers(0, 256, (40, 64, 3), dtype=np.uint8)
    want = np.ascontiguousarray(orc.colourspace(a, "lab", "srgb")[8:24, 16:48])
    out = np.zeros_like(want)
    src = np.ascontiguousarray(a[4:30, 10:60])
    rin = region(vb, src, 10, 4, interp=22)
    rout = region(vb, out, 16, 8, interp=13)
    vb._check(vb.lib().vb200_colour_gen(C.byref(rout), C.byref(rin), 13))
    assert np.array_equal(out, want)
