"""Generate tests/golden/*.npz from the REFERENCE's own code (oracle/_ref, i.e.
libvips' sources compiled in place under the GLib-free shim).  Run in the
build container where /root/reference exists:

    python tests/golden/make_golden.py

Inputs are regenerated from numpy seeds by the tests (cases.py), so only the
expected outputs are stored.  /root/reference is NOT needed to run the tests.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from cases import resample_cases, seam_cases, make_input  # noqa: E402
from oracle import pyref  # noqa: E402


def run_ref(case):
    a = make_input(case)
    op = case["op"]
    if op == "thumbnail":
        return pyref.thumbnail_image(a, case["width"], case.get("height"), case.get("size", "both"))
    r = pyref.RefImage.from_array(a)
    if op == "shrinkv":
        return r.shrinkv(case["f"], case.get("ceil", False)).numpy()
    if op == "shrinkh":
        return r.shrinkh(case["f"], case.get("ceil", False)).numpy()
    if op == "reducev":
        return r.reducev(case["f"], case.get("kernel", "lanczos3"), case.get("gap", 0.0)).numpy()
    if op == "reduceh":
        return r.reduceh(case["f"], case.get("kernel", "lanczos3"), case.get("gap", 0.0)).numpy()
    if op == "resize":
        return r.resize(case["scale"], case.get("vscale"), case.get("kernel", "lanczos3"), case.get("gap", 2.0)).numpy()
    if op == "premultiply":
        return r.premultiply(case.get("max_alpha", 0.0), case.get("uchar", False)).numpy()
    if op == "unpremultiply":
        return r.unpremultiply(case.get("max_alpha", 0.0), case.get("uchar", False)).numpy()
    raise ValueError(op)


def main():
    out = {}
    for case in resample_cases():
        out[case["name"]] = run_ref(case)
    path = os.path.join(HERE, "resample_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d cases, %d bytes" % (path, len(out), os.path.getsize(path)))

    # tiled evaluations for the seam tests: the reference's generate() called with explicit sink tiles
    out = {}
    for case in seam_cases():
        r = pyref.RefImage.from_array(make_input(case))
        r = r.reducev(case["f"]) if case["op"] == "reducev" else r.reduceh(case["f"])
        tw, th = case["tile"]
        out[case["name"]] = r.numpy(tile=(tw if tw else r.shape[1], th))
    path = os.path.join(HERE, "seams_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d cases, %d bytes" % (path, len(out), os.path.getsize(path)))

    # the one stored-answer fixture the reference's test-suite has on this path
    # (test/test-suite/test_resample.py:234-238): rgba.png -> thumbnail(64, linear=True)
    # .flatten(255).avg() must be within 1 of rgba-correct.ppm.  Decoded here with PIL.
    from PIL import Image
    images = "/root/reference/test/test-suite/images"
    rgba = np.array(Image.open(os.path.join(images, "rgba.png")))
    correct = np.array(Image.open(os.path.join(images, "rgba-correct.ppm")))
    path = os.path.join(HERE, "rgba_fixture.npz")
    np.savez_compressed(path, rgba=rgba, correct_avg=np.float64(correct.mean()), correct_shape=np.array(correct.shape))
    print("wrote %s: %d bytes" % (path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
