"""The golden cases: name, seeded input recipe, op and arguments.  Shared by
make_golden.py (which runs the reference) and tests/test_golden.py."""
import numpy as np

DT = {"u8": np.uint8, "i8": np.int8, "u16": np.uint16, "i16": np.int16, "u32": np.uint32, "i32": np.int32,
      "f32": np.float32}


def make_input(case):
    rng = np.random.default_rng(case["seed"])
    h, w, b = case["shape"]
    dt = np.dtype(DT[case["dtype"]])
    if dt.kind == "f":
        return (rng.random((h, w, b)) * 255).astype(dt)
    info = np.iinfo(dt)
    return rng.integers(info.min, int(info.max) + 1, (h, w, b), dtype=np.int64).astype(dt)


def resample_cases():
    cases = []
    seed = 1000

    def add(**kw):
        nonlocal seed
        seed += 1
        kw["seed"] = seed
        kw["name"] = "%03d_%s_%s" % (len(cases), kw["op"], kw["dtype"])
        cases.append(kw)

    for dt in DT:
        add(op="shrinkv", dtype=dt, shape=(37, 29, 3), f=3)
        add(op="shrinkh", dtype=dt, shape=(23, 41, 4), f=4, ceil=True)
        add(op="reducev", dtype=dt, shape=(61, 33, 3), f=1.7)
        add(op="reduceh", dtype=dt, shape=(33, 61, 3), f=2.3, kernel="cubic")
        add(op="resize", dtype=dt, shape=(96, 120, 3), scale=0.23)
    for k in ("nearest", "linear", "cubic", "mitchell", "lanczos2", "lanczos3", "mks2013", "mks2021"):
        add(op="reducev", dtype="u8", shape=(64, 40, 4), f=2.0, kernel=k)
        add(op="reduceh", dtype="u8", shape=(40, 64, 4), f=1.5, kernel=k)
    add(op="reducev", dtype="u8", shape=(200, 24, 4), f=8.0)  # 49 taps
    add(op="reducev", dtype="u8", shape=(200, 24, 4), f=8.0, gap=2.0)
    add(op="reduceh", dtype="u8", shape=(24, 200, 4), f=8.0, gap=2.0)
    for dt in ("u8", "u16", "i16", "f32"):
        add(op="premultiply", dtype=dt, shape=(17, 19, 4))
        add(op="unpremultiply", dtype=dt, shape=(17, 19, 4))
    add(op="premultiply", dtype="u8", shape=(64, 64, 4), uchar=True)
    add(op="unpremultiply", dtype="u8", shape=(64, 64, 4), uchar=True)
    add(op="thumbnail", dtype="u8", shape=(512, 512, 4), width=64)  # shrink 8: the BASELINE chain
    add(op="thumbnail", dtype="u8", shape=(301, 517, 4), width=50)
    add(op="thumbnail", dtype="u8", shape=(400, 300, 3), width=60)
    add(op="thumbnail", dtype="u8", shape=(256, 384, 4), width=48, height=100, size="force")
    add(op="thumbnail", dtype="u8", shape=(300, 300, 1), width=75)
    return cases


def seam_cases():
    """Tiled evaluations for the generate()-shaped / scanline seams (tests/test_seams_gpu.py): the same
    image pulled through the reference's generate() with explicit sink tiles, so that the rects --
    and with a non-representable factor such as 1.7 the per-rect coordinate stepping -- are pinned."""
    out = []
    for op, shape in (("reducev", (333, 300, 4)), ("reduceh", (150, 420, 4)), ("reducev", (260, 150, 3)),
                      ("reduceh", (70, 311, 3))):
        for tname, tile in (("smalltile", (128, 128)), ("fatstrip", (0, 16))):
            out.append({"op": op, "dtype": "u8", "shape": shape, "f": 1.7, "tile": tile, "seed": 4242 + len(out),
                        "name": "seam%02d_%s_%s_%dband" % (len(out), op, tname, shape[2])})
    return out
