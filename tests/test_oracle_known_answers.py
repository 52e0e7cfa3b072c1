"""The known answers the reference's own test-suite holds for the resample path, restated on the ORACLE (CPU):
test/test-suite/test_resample.py:77-146.  The GPU suite asks the same of the CUDA path (tests/test_resample_gpu.py); the
colour, convolution, morphology and conversion known answers live next to their ops (tests/test_colour.py,
test_convolution.py, test_morphology.py, test_widen_*.py)."""
import numpy as np
import pytest

from oracle import pyoracle as orc

KERNELS = ["nearest", "linear", "cubic", "lanczos2", "lanczos3", "mks2013", "mks2021"]
FORMATS = [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.float32, np.float64]


def photo(h=221, w=145):
    """a smooth synthetic stand-in for sample.jpg, cast down to 0 .. 127 as the reference test does (test_resample.py:79-81)"""
    yy, xx = np.mgrid[0:h, 0:w]
    a = np.stack([64 + 50 * np.sin(xx / 17.0 + yy / 31.0), 64 + 45 * np.cos(xx / 23.0 - yy / 13.0), (xx + 2 * yy) % 128], -1)
    rng = np.random.default_rng(7)
    return np.clip(a + rng.normal(0, 4, a.shape), 0, 127).astype(np.int8)


@pytest.mark.parametrize("kernel", KERNELS)
def test_reduce_keeps_the_average(kernel):
    """test_resample.py:83-91: |avg(reduce(x)) - avg(x)| < 2 for every format, kernel and factor"""
    im = photo()
    for fac in (1, 1.1, 1.5, 1.999):
        for fmt in FORMATS:
            x = im.astype(fmt)
            v = orc.reducev(x, fac, kernel, 0.0)
            r = orc.reduceh(v, fac, kernel, 0.0)
            assert abs(r.astype(np.float64).mean() - im.astype(np.float64).mean()) < 2, (fac, fmt, kernel)


def test_constant_images_survive_reduce():
    """test_resample.py:93-103"""
    for const in (0, 1, 2, 254, 255):
        im = np.full((10, 10, 1), const, np.uint8)
        for kernel in KERNELS:
            shr = orc.reduceh(orc.reducev(im, 2, kernel, 0.0), 2, kernel, 0.0)
            assert shr.shape == (5, 5, 1) and (shr == const).all(), (const, kernel)


def test_reduceh_nearest_of_two_pixels():
    """test_resample.py:105-111, libvips issue 4864: a 2 x 2 image under reduceh(1.5, nearest) is one pixel wide"""
    im = np.array([[[255, 0, 0], [0, 255, 0]], [[0, 0, 255], [255, 255, 0]]], np.uint8)
    assert orc.reduceh(im, 1.5, "nearest", 0.0).shape == (2, 1, 3)


def test_resize_geometry():
    """test_resample.py:113-131: a quarter rounds to nearest; 100 x 1 -> 50 x 1; 1600 x 1000 at 10 / 1600 -> 10 x 6 (double
    precision in reduce{h,v})"""
    im = photo(442, 290).astype(np.uint8)
    r = orc.resize(im, 0.25)
    assert r.shape[:2] == (int(442 / 4.0 + 0.5), int(290 / 4.0 + 0.5))
    assert orc.resize(np.zeros((1, 100, 1), np.uint8), 0.5).shape == (1, 50, 1)
    assert orc.resize(np.zeros((1000, 1600, 1), np.uint8), 10.0 / 1600).shape == (6, 10, 1)


@pytest.mark.parametrize("scale", [8, 9.4, 16])
def test_resize_keeps_the_edges(scale):
    """test_resample.py:133-146: a black image with a one-pixel red border: the mid-edge pixels of the result are not black
    (the round-up option of shrink)"""
    im = np.zeros((2047, 2049, 3), np.uint8)
    im[0, :, 0] = im[-1, :, 0] = 255
    im[:, 0, 0] = im[:, -1, 0] = 255
    x = orc.resize(im, 1 / scale, 1 / scale)
    h, w = x.shape[:2]
    for (px, py) in ((round(w / 2), 0), (w - 1, round(h / 2)), (round(w / 2), h - 1), (0, round(h / 2))):
        assert x[py, px, 0] != 0, (scale, px, py)
