"""linear=TRUE thumbnails: sRGB -> scRGB, float premultiply / resize /
unpremultiply, scRGB -> sRGB.  Known answer from the reference's test-suite
(test_resample.py:234-238) + GPU parity against the oracle."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = np.load(os.path.join(HERE, "golden", "rgba_fixture.npz"))


def flatten_avg(t):
    t = t.astype(np.float64)
    al = t[..., 3:4] / 255.0
    return float((t[..., :3] * al + 255.0 * (1 - al)).mean())


def test_oracle_rgba_correct_known_answer(oracle):
    t = oracle.thumbnail_image(FIX["rgba"], 64, linear=True)
    assert t.shape[:2] == tuple(FIX["correct_shape"][:2])
    assert abs(flatten_avg(t) - float(FIX["correct_avg"])) < 1


@pytest.mark.gpu
def test_gpu_rgba_correct_known_answer(vb, oracle):
    t = vb.Image(FIX["rgba"]).thumbnail_image(64, linear=True).numpy()
    assert abs(flatten_avg(t) - float(FIX["correct_avg"])) < 1
    assert np.array_equal(t, oracle.thumbnail_image(FIX["rgba"], 64, linear=True))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,target,bands", [((512, 512), 64, 4), ((301, 517), 50, 4), ((400, 300), 60, 3),
                                                 ((640, 480), 111, 4)])
def test_gpu_linear_thumbnail_parity(vb, oracle, shape, target, bands):
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, shape + (bands,), dtype=np.uint8)
    got = vb.Image(a).thumbnail_image(target, linear=True).numpy()
    want = oracle.thumbnail_image(a, target, linear=True)
    assert got.shape == want.shape and np.array_equal(got, want)
