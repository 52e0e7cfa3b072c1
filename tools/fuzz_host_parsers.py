"""tools/fuzz_host_parsers.py -- mutation fuzzing of the host-side parsers that read untrusted bytes (the JPEG marker
parser + the decoder's host twin, the ICC profile parser), meant to run against an AddressSanitizer build:

    VB200_LIB=/tmp/asan/libvb200_asan.so LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \
        python tools/fuzz_host_parsers.py [seconds]

Every call must either succeed or fail with vb.Error; a crash or an ASan report is a bug.  No GPU is used."""
import io
import sys
import time

import numpy as np
from PIL import Image as PIL

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0] + "/tests")
import libvips_b200 as vb  # noqa: E402


def jpegs(rng):
    out = []
    for shape in ((33, 47, 3), (64, 64, 3), (17, 90, 1)):
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        if shape[2] == 1:
            a = a[:, :, 0]
        for kw in (dict(), dict(subsampling=0), dict(subsampling=1), dict(progressive=True), dict(restart_marker_blocks=2),
                   dict(progressive=True, restart_marker_blocks=3), dict(optimize=True)):
            if a.ndim == 2:
                kw = {k: v for k, v in kw.items() if k != "subsampling"}
            b = io.BytesIO()
            PIL.fromarray(a).save(b, "JPEG", quality=int(rng.integers(30, 96)), **kw)
            out.append(b.getvalue())
    return out


def mutate(rng, s):
    s = bytearray(s)
    k = rng.integers(0, 6)
    if k == 0:
        return bytes(s[: rng.integers(0, len(s))])
    if k == 1:
        for _ in range(rng.integers(1, 8)):
            s[rng.integers(0, len(s))] ^= 1 << rng.integers(0, 8)
    elif k == 2:  # damage a marker segment length
        pos = [i for i in range(len(s) - 3) if s[i] == 0xFF and 0xC0 <= s[i + 1] <= 0xFE]
        if pos:
            i = pos[rng.integers(0, len(pos))]
            s[i + 2] = rng.integers(0, 256)
            s[i + 3] = rng.integers(0, 256)
    elif k == 3:
        i = rng.integers(0, len(s))
        s[i:i] = bytes(rng.integers(0, 256, rng.integers(1, 40), dtype=np.uint8))
    elif k == 4:
        i = rng.integers(0, len(s))
        del s[i:i + rng.integers(1, 60)]
    else:
        i = rng.integers(0, len(s))
        n = rng.integers(1, 30)
        s[i:i + n] = bytes(rng.integers(0, 256, n, dtype=np.uint8))
    return bytes(s)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(time.time()))
    good = jpegs(rng)
    import icc_fixtures as F
    import test_icc as T
    profiles = [F.rgb_profile(t) for t in ("srgb", "gamma", "para4", "table")] + [F.grey_profile(), F.ink_profile(),
                                                                               F.lut_v4_rgb_profile("XYZ "), F.lut_v4_rgb_profile("Lab ")]
    profiles = [bytes(p) for p in profiles]
    t0 = time.time()
    n = fails = ok = 0
    px = rng.integers(0, 256, (16, 3), dtype=np.uint8)
    while time.time() - t0 < budget:
        s = mutate(rng, good[rng.integers(0, len(good))])
        for shrink in (1, 2, 8):
            try:
                vb.jpeg_decode_host_twin(s, shrink)
                ok += 1
            except vb.Error:
                fails += 1
        if profiles:
            p = mutate(rng, profiles[rng.integers(0, len(profiles))])
            for mode in (0, 1):
                try:
                    T.host_eval(mode, px if mode == 0 else px.astype(np.float32), p, intent=int(rng.integers(0, 4)), pcs=int(rng.integers(0, 2)))
                    ok += 1
                except vb.Error:
                    fails += 1
        n += 1
    print("fuzz: %d mutants, %d calls ok, %d rejected, %d profiles, no crash" % (n, ok, fails, len(profiles)))


if __name__ == "__main__":
    main()
