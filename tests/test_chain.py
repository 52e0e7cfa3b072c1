"""The chain pump for unfused graphs (SURVEY 8f rank 2, vb200_chain_*): a list of operations run over a
batch of host images with every intermediate on the device.  Expected pixels: the oracle's operations
applied one after the other (each pinned to the reference by its own test file)."""
import numpy as np
import pytest

from oracle import pyconv, pyoracle as orc

pytestmark = pytest.mark.gpu


def test_chain_matches_oracle_op_by_op(vb):
    rng = np.random.default_rng(90)
    imgs = [rng.integers(0, 256, s, dtype=np.uint8) for s in ((300, 400, 3), (257, 129, 3), (64, 64, 3), (500, 333, 3))]
    chain = vb.Chain().resize(0.4).gaussblur(1.2, 0.2, "integer").sharpen().colourspace("lab")
    got = chain.run(imgs)
    for a, g in zip(imgs, got):
        want = orc.resize(a, 0.4)
        want = pyconv.gaussblur(want, 1.2, 0.2, "integer")
        want = pyconv.sharpen(want, "srgb")
        want = orc.colourspace(want, "lab", "srgb")
        assert g.numpy().shape == want.shape and np.array_equal(g.numpy(), want)
    # the same chain again (streams and pool are reused), and a different batch size
    again = chain.run(imgs[:2])
    assert all(np.array_equal(x.numpy(), y.numpy()) for x, y in zip(again, got))


def test_chain_equals_standalone_calls(vb):
    rng = np.random.default_rng(91)
    a = rng.integers(0, 256, (240, 320, 4), dtype=np.uint8)
    mask = np.array([[1.0, 2.0, 1.0], [2.0, 4.0, 2.0], [1.0, 2.0, 1.0]])
    got = vb.Chain().premultiply(uchar=True).reduce(2.5, 1.7).unpremultiply(uchar=True).conv(mask, 16.0, 0.0, "integer").run([a])[0]
    step = vb.Image(a).premultiply(uchar=True).reduce(2.5, 1.7).unpremultiply(uchar=True).conv(mask, 16.0, 0.0, "integer")
    assert np.array_equal(got.numpy(), step.numpy())
    # float convsep with a column mask in a chain
    col = np.array([[1.0], [3.0], [5.0], [2.0]])
    got = vb.Chain().convsep(col, 11.0, 1.0, "float").run([a])[0]
    assert np.array_equal(got.numpy(), pyconv.convsep(a, col, 11.0, 1.0, "float"))


def test_chain_errors_surface(vb):
    a = np.zeros((32, 32, 3), np.uint8)
    with pytest.raises(vb.Error, match="reduce factor should be >= 1.0"):
        vb.Chain().reduce(0.5, 2.0).run([a])
    with pytest.raises(vb.Error):
        vb.Chain().colourspace("cmyk").run([a])
