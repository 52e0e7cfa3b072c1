"""Golden vectors produced by the reference's own code (tests/golden/make_golden.py).

CPU: the oracle must reproduce them.  GPU: the CUDA path must reproduce them.
Neither needs /root/reference at run time.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from cases import make_input, resample_cases  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "resample_ref.npz"))
CASES = resample_cases()


def run_oracle(orc, c):
    a = make_input(c)
    op = c["op"]
    if op == "thumbnail":
        return orc.thumbnail_image(a, c["width"], c.get("height"), c.get("size", "both"))
    if op == "shrinkv":
        return orc.shrinkv(a, c["f"], c.get("ceil", False))
    if op == "shrinkh":
        return orc.shrinkh(a, c["f"], c.get("ceil", False))
    if op == "reducev":
        gap = c.get("gap", 0.0)
        return orc.reducev(a, c["f"], c.get("kernel", "lanczos3"), gap, rect_h=128 if gap else 16)
    if op == "reduceh":
        return orc.reduceh(a, c["f"], c.get("kernel", "lanczos3"), c.get("gap", 0.0))
    if op == "resize":
        return orc.resize(a, c["scale"], c.get("vscale"), c.get("kernel", "lanczos3"), c.get("gap", 2.0))
    if op == "premultiply":
        return orc.premultiply(a, 255.0, c.get("uchar", False))
    if op == "unpremultiply":
        return orc.unpremultiply(a, 255.0, c.get("uchar", False))
    raise ValueError(op)


def run_gpu(vb, c):
    im = vb.Image(make_input(c))
    op = c["op"]
    if op == "thumbnail":
        return im.thumbnail_image(c["width"], c.get("height"), c.get("size", "both")).numpy()
    if op == "shrinkv":
        return im.shrinkv(c["f"], c.get("ceil", False)).numpy()
    if op == "shrinkh":
        return im.shrinkh(c["f"], c.get("ceil", False)).numpy()
    if op == "reducev":
        return im.reducev(c["f"], c.get("kernel", "lanczos3"), c.get("gap", 0.0)).numpy()
    if op == "reduceh":
        return im.reduceh(c["f"], c.get("kernel", "lanczos3"), c.get("gap", 0.0)).numpy()
    if op == "resize":
        return im.resize(c["scale"], c.get("vscale"), c.get("kernel", "lanczos3"), c.get("gap", 2.0)).numpy()
    if op == "premultiply":
        return im.premultiply(uchar=c.get("uchar", False)).numpy()
    if op == "unpremultiply":
        return im.unpremultiply(uchar=c.get("uchar", False)).numpy()
    raise ValueError(op)


def test_fixture_is_complete():
    assert sorted(GOLD.files) == sorted(c["name"] for c in CASES)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_reference_golden(oracle, case):
    got = run_oracle(oracle, case)
    want = GOLD[case["name"]]
    assert got.shape == want.shape and got.dtype == want.dtype
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_gpu_reproduces_reference_golden(vb, case):
    got = run_gpu(vb, case)
    want = GOLD[case["name"]]
    assert got.shape == want.shape and got.dtype == want.dtype
    assert np.array_equal(got, want)
