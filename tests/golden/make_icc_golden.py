"""Generate tests/golden/icc_lcms.npz: outputs of lcms2 itself (the 2.18 inside Pillow, through oracle/pylcms.py,
which makes the reference's own calls: icc_transform.c:459, :931, :1094, :1219) on seeded inputs and the synthetic
profiles of tests/icc_fixtures.py.  tests/test_icc.py::test_gpu_against_lcms2_fixtures compares the CUDA kernel
with them, so the GPU suite holds an lcms2 comparison that needs neither lcms2 nor /root/reference at run time.

    python tests/golden/make_icc_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import icc_fixtures as F  # noqa: E402
from oracle import pylcms  # noqa: E402


def inputs():
    """the seeded inputs, shared with the test"""
    rng = np.random.default_rng(2024)
    return {
        "rgb8": rng.integers(0, 256, (6000, 3), dtype=np.uint8),
        "rgbf": rng.random((6000, 3), dtype=np.float32),
        "rgb16": rng.integers(0, 65536, (6000, 3), dtype=np.uint16),
        "cmyk8": rng.integers(0, 256, (6000, 4), dtype=np.uint8),
        "grey8": np.arange(256, dtype=np.uint8).reshape(-1, 1),
    }


def main():
    assert pylcms.available(), "no lcms2 next to Pillow"
    I = inputs()
    rgb, gamma, table, grey, ink = F.rgb_profile("srgb"), F.rgb_profile("gamma"), F.rgb_profile("table"), F.grey_profile(), F.ink_profile()
    v4 = F.lut_v4_rgb_profile("Lab ")
    out = {}
    out["import_rgb8_lab"] = pylcms.icc_import(I["rgb8"], rgb)
    out["import_rgbf_lab"] = pylcms.icc_import(I["rgbf"], rgb)
    out["import_rgb8_xyz"] = pylcms.icc_import(I["rgb8"], rgb, pcs="xyz")
    out["import_table8_lab"] = pylcms.icc_import(I["rgb8"], table)
    lab = out["import_rgb8_lab"]
    out["export_lab_rgb8"] = pylcms.icc_export(lab, rgb)
    out["export_lab_rgb16"] = pylcms.icc_export(lab, rgb, depth=16)
    out["export_lab_gamma8"] = pylcms.icc_export(lab, gamma)
    out["transform_rgb8_gamma8"] = pylcms.icc_transform(I["rgb8"], rgb, gamma)
    out["transform_rgb16_gamma16"] = pylcms.icc_transform(I["rgb16"], rgb, gamma, depth=16)
    out["import_grey8_lab"] = pylcms.icc_import(I["grey8"], grey)
    out["export_lab_grey8"] = pylcms.icc_export(out["import_grey8_lab"], grey)
    out["import_cmyk8_lab"] = pylcms.icc_import(I["cmyk8"], ink)
    out["export_lab_cmyk8"] = pylcms.icc_export(lab, ink)
    out["import_v4_rgb8_lab"] = pylcms.icc_import(I["rgb8"], v4)
    out["export_lab_v4_rgb8"] = pylcms.icc_export(lab, v4)
    path = os.path.join(HERE, "icc_lcms.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d cases, %d bytes" % (path, len(out), os.path.getsize(path)))


if __name__ == "__main__":
    main()
