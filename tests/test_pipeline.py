"""BASELINE config 5: the batched thumbnail + sharpen + sRGB stream, and the fused sharpen kernel on its
own.  Expected pixels: the oracle's vips_sharpen (pinned bit for bit to the reference's sharpen.c under
oracle/_ref by tests/test_convolution.py::test_sharpen_against_reference_sharpen_c) over the oracle's
vips_thumbnail_image."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyconv, pyoracle as orc

pytestmark = pytest.mark.gpu


def rnd(rng, shape):
    return rng.integers(0, 256, shape, dtype=np.uint8)


def photo(rng, h, w, b):
    """smooth content with texture: small L differences exercise the LUT's centre and both slopes"""
    y, x = np.mgrid[0:h, 0:w]
    base = 128 + 90 * np.sin(x / 17.0) * np.cos(y / 23.0)
    img = base[:, :, None] + rng.normal(0, 6, (h, w, b)) + np.array([10, -20, 30, 0][:b])
    out = np.clip(img, 0, 255).astype(np.uint8)
    if b == 4:
        out[:, :, 3] = rng.integers(0, 256, (h, w), dtype=np.uint8)
    return out


@pytest.mark.parametrize("bands", [3, 4])
@pytest.mark.parametrize("kw", [{}, {"sigma": 1.5, "m2": 5.0, "y2": 20.0}, {"sigma": 0.8, "x1": 1.0, "m1": 0.5, "y3": 5.0},
                                {"sigma": 3.0}])
def test_fused_sharpen_matches_oracle(vb, bands, kw):
    rng = np.random.default_rng(50 + bands)
    for a in (rnd(rng, (97, 131, bands)), photo(rng, 64, 200, bands), rnd(rng, (1, 1, bands)), rnd(rng, (33, 2, bands))):
        got = vb.Image(a, "srgb").sharpen(**kw).numpy()
        assert np.array_equal(got, pyconv.sharpen(a, "srgb", **kw)), (a.shape, kw)


def test_fused_sharpen_is_the_path_taken(vb):
    a = rnd(np.random.default_rng(1), (64, 64, 3))
    n0 = vb.launch_count()
    vb.Image(a, "srgb").sharpen()
    assert vb.launch_count() - n0 == 1, "8-bit sRGB sharpen must be the single fused kernel"


def test_sharpen_batch_device(vb):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(60)
    frames = np.stack([photo(rng, 96, 80, 4) for _ in range(5)])
    d = torch.from_numpy(frames).cuda()
    o = torch.empty_like(d)
    vb.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        L = vb.lib()
        vb._check(L.vb200_sharpen_batch_device(d.data_ptr(), 96 * 80 * 4, o.data_ptr(), 96 * 80 * 4, 5, 80, 96, 4, 0.5, 2.0, 10.0,
                                               20.0, 0.0, 3.0))
        torch.cuda.synchronize()
        got = o.cpu().numpy()
        for i in range(5):
            assert np.array_equal(got[i], pyconv.sharpen(frames[i], "srgb")), i
        # in-place is refused: tiles read their neighbours' halo
        assert L.vb200_sharpen_batch_device(d.data_ptr(), 96 * 80 * 4, d.data_ptr(), 96 * 80 * 4, 5, 80, 96, 4, 0.5, 2.0, 10.0, 20.0,
                                            0.0, 3.0) == -1
        assert b"overlap" in L.vb200_error_buffer()
        L.vb200_error_clear()
    finally:
        vb.set_stream(0)


@pytest.mark.parametrize("shape,target", [((512, 512, 4), 64), ((600, 450, 4), 120), ((400, 300, 3), 60), ((1024, 768, 4), 128)])
def test_thumbnail_sharpen_pipeline(vb, shape, target):
    """config 5 on host frames through the pump, and on device frames"""
    rng = np.random.default_rng(70)
    h, w, b = shape
    frames = np.stack([photo(rng, h, w, b), rnd(rng, shape), photo(rng, h, w, b)[::-1].copy()])
    plan = vb.ThumbnailPlan(w, h, b, target)
    plain = plan.run_host(frames)
    plan.set_sharpen()
    got = plan.run_host(frames)
    for i in range(len(frames)):
        thumb = orc.thumbnail_image(frames[i], target)
        assert np.array_equal(plain[i], thumb)
        assert np.array_equal(got[i], pyconv.sharpen(thumb, "srgb")), i
    plan.set_sharpen(sigma=1.2, m2=4.0)
    got = plan.run_host(frames[:1])
    assert np.array_equal(got[0], pyconv.sharpen(orc.thumbnail_image(frames[0], target), "srgb", sigma=1.2, m2=4.0))
    plan.set_sharpen(sigma=0)  # off again
    assert np.array_equal(plan.run_host(frames[:1])[0], plain[0])
