"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline /
--impl reference leg and __graft_entry__.smoke().  Never by libvips_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FMT = {np.dtype(np.uint8): 0, np.dtype(np.int8): 1, np.dtype(np.uint16): 2, np.dtype(np.int16): 3,
       np.dtype(np.uint32): 4, np.dtype(np.int32): 5, np.dtype(np.float32): 6, np.dtype(np.float64): 8}
DTYPE = {v: k for k, v in FMT.items()}

KERNELS = {"nearest": 0, "linear": 1, "cubic": 2, "mitchell": 3, "lanczos2": 4, "lanczos3": 5,
           "mks2013": 6, "mks2021": 7}
SIZES = {"both": 0, "up": 1, "down": 2, "force": 3}


class ReduceGeom(C.Structure):
    _fields_ = [("in_size", C.c_int), ("out_size", C.c_int), ("int_shrink", C.c_int),
                ("shrunk_size", C.c_int), ("residual", C.c_double), ("n_point", C.c_int),
                ("offset", C.c_double)]


def build(force=False):
    """Compile the oracle with the flags its header documents."""
    subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []) + ["liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_reduce_make_mask.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        _LIB.orc_reduce_get_points.argtypes = [C.c_int, C.c_double]
        _LIB.orc_reduce_geometry.argtypes = [C.c_int, C.c_double, C.c_int, C.c_double, C.POINTER(ReduceGeom)]
        _LIB.orc_reduce_tables.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    return _LIB


def _img(a):
    a = np.ascontiguousarray(a)
    if a.ndim == 2:
        a = a[:, :, None]
    return a, a.shape[0], a.shape[1], a.shape[2], FMT[a.dtype]


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _k(kernel):
    return KERNELS[kernel] if isinstance(kernel, str) else int(kernel)


def reduce_get_points(kernel, shrink):
    return lib().orc_reduce_get_points(_k(kernel), C.c_double(shrink))


def reduce_make_mask(kernel, n_point, shrink, x):
    c = np.zeros(n_point, np.float64)
    lib().orc_reduce_make_mask(_p(c), _k(kernel), n_point, shrink, x)
    return c


def reduce_tables(kernel, n_point, residual):
    f = np.zeros((65, n_point), np.float64)
    s = np.zeros((65, n_point), np.int16)
    lib().orc_reduce_tables(_k(kernel), n_point, residual, _p(f), _p(s))
    return f, s


def reduce_geometry(in_size, shrink, kernel="lanczos3", gap=0.0):
    g = ReduceGeom()
    if lib().orc_reduce_geometry(in_size, shrink, _k(kernel), gap, C.byref(g)):
        raise ValueError("bad reduce geometry")
    return g


def shrinkv(a, vshrink, ceil=False):
    a, h, w, b, f = _img(a)
    oh = h if vshrink == 1 else lib().orc_shrink_size(h, vshrink, int(ceil))
    out = np.empty((oh, w, b), a.dtype)
    if lib().orc_shrinkv(_p(a), w, h, b, f, vshrink, int(ceil), _p(out)):
        raise ValueError("shrinkv")
    return out


def shrinkh(a, hshrink, ceil=False):
    a, h, w, b, f = _img(a)
    ow = w if hshrink == 1 else lib().orc_shrink_size(w, hshrink, int(ceil))
    out = np.empty((h, ow, b), a.dtype)
    if lib().orc_shrinkh(_p(a), w, h, b, f, hshrink, int(ceil), _p(out)):
        raise ValueError("shrinkh")
    return out


def reducev(a, vshrink, kernel="lanczos3", gap=0.0, rect_h=0):
    a, h, w, b, f = _img(a)
    g = reduce_geometry(h, vshrink, kernel, gap)
    out = np.empty((g.out_size, w, b), a.dtype)
    if lib().orc_reducev(_p(a), w, h, b, f, C.c_double(vshrink), _k(kernel), C.c_double(gap), rect_h, _p(out)):
        raise ValueError("reducev")
    return out


def reduceh(a, hshrink, kernel="lanczos3", gap=0.0, rect_w=0):
    a, h, w, b, f = _img(a)
    g = reduce_geometry(w, hshrink, kernel, gap)
    out = np.empty((h, g.out_size, b), a.dtype)
    if lib().orc_reduceh(_p(a), w, h, b, f, C.c_double(hshrink), _k(kernel), C.c_double(gap), rect_w, _p(out)):
        raise ValueError("reduceh")
    return out


def _pre_dtype(dt, uchar):
    if uchar and dt == np.uint8:
        return np.uint8
    return np.float64 if dt == np.float64 else np.float32


def premultiply(a, max_alpha=255.0, uchar=False):
    a, h, w, b, f = _img(a)
    out = np.empty((h, w, b), a.dtype if b == 1 else _pre_dtype(a.dtype, uchar))
    if lib().orc_premultiply(_p(a), w, h, b, f, C.c_double(max_alpha), int(uchar), _p(out)):
        raise ValueError("premultiply")
    return out


def unpremultiply(a, max_alpha=255.0, uchar=False):
    a, h, w, b, f = _img(a)
    out = np.empty((h, w, b), a.dtype if b == 1 else _pre_dtype(a.dtype, uchar))
    if lib().orc_unpremultiply(_p(a), w, h, b, f, C.c_double(max_alpha), int(uchar), _p(out)):
        raise ValueError("unpremultiply")
    return out


def resize(a, scale, vscale=None, kernel="lanczos3", gap=2.0, tile=(0, 0)):
    a, h, w, b, f = _img(a)
    vscale = scale if vscale is None else vscale
    ow, oh = C.c_int(), C.c_int()
    if lib().orc_resize_size(w, h, C.c_double(scale), C.c_double(vscale), _k(kernel), C.c_double(gap),
                             C.byref(ow), C.byref(oh)):
        raise ValueError("resize size")
    out = np.empty((oh.value, ow.value, b), a.dtype)
    if lib().orc_resize(_p(a), w, h, b, f, C.c_double(scale), C.c_double(vscale), _k(kernel), C.c_double(gap),
                        tile[0], tile[1], _p(out)):
        raise ValueError("resize")
    return out


def thumbnail_size(w, h, width, height=None, size="both"):
    height = width if height is None else height
    hs, vs, ow, oh = C.c_double(), C.c_double(), C.c_int(), C.c_int()
    if lib().orc_thumbnail_size(w, h, width, height, SIZES[size], C.byref(hs), C.byref(vs), C.byref(ow),
                                C.byref(oh)):
        raise ValueError("thumbnail size")
    return hs.value, vs.value, ow.value, oh.value


def thumbnail_image(a, width, height=None, size="both", has_alpha=None, tile=(0, 0), linear=False):
    a, h, w, b, f = _img(a)
    assert a.dtype == np.uint8
    if has_alpha is None:
        has_alpha = b == 2 or b >= 4  # vips_image_hasalpha for B_W (< 3 bands) / sRGB (3+), image.c:3113-3119
    height = width if height is None else height
    _, _, ow, oh = thumbnail_size(w, h, width, height, size)
    out = np.empty((oh, ow, b), np.uint8)
    fn = lib().orc_thumbnail_image_linear if linear else lib().orc_thumbnail_image
    if fn(_p(a), w, h, b, width, height, SIZES[size], int(has_alpha), tile[0], tile[1], _p(out)):
        raise ValueError("thumbnail_image")
    return out


# ------------------------------------------------------------------ colour
STEPS = {"sRGB2scRGB": 1, "scRGB2XYZ": 2, "XYZ2Lab": 3, "Lab2LabS": 4, "LabS2Lab": 5, "Lab2XYZ": 6,
         "XYZ2scRGB": 7, "scRGB2sRGB": 8, "scRGB2RGB16": 9, "RGB162scRGB": 10, "Lab2LCh": 11, "LCh2Lab": 12,
         "XYZ2Yxy": 13, "Yxy2XYZ": 14, "sRGB2RGB16": 15, "RGB162sRGB": 16, "sRGB2HSV": 17, "HSV2sRGB": 18,
         "scRGB2BW": 19, "scRGB2BW16": 20, "BW2sRGB": 21, "GREY162RGB16": 22}
SPACES = {"xyz": 12, "lab": 13, "lch": 19, "labs": 21, "srgb": 22, "yxy": 23, "rgb16": 25, "scrgb": 28, "b-w": 1,
          "multiband": 0, "grey16": 26, "hsv": 29}
# (input dtype the step wants, output dtype, output interpretation)
STEP_IO = {1: (np.uint8, np.float32, 28), 10: (np.uint16, np.float32, 28), 2: (np.float32, np.float32, 12),
           3: (np.float32, np.float32, 13), 4: (np.float32, np.int16, 21), 5: (np.int16, np.float32, 13),
           6: (np.float32, np.float32, 12), 7: (np.float32, np.float32, 28), 8: (np.float32, np.uint8, 22),
           9: (np.float32, np.uint16, 25), 11: (np.float32, np.float32, 19), 12: (np.float32, np.float32, 13),
           13: (np.float32, np.float32, 23), 14: (np.float32, np.float32, 12), 15: (np.uint8, np.uint16, 25),
           16: (np.uint16, np.uint8, 22), 17: (np.uint8, np.uint8, 29), 18: (np.uint8, np.uint8, 22),
           19: (np.float32, np.uint8, 1), 20: (np.float32, np.uint16, 26), 21: (None, None, 22), 22: (None, None, 25)}
# bands the step adds (BW2sRGB / GREY162RGB16: one band becomes three) or removes (scRGB2BW: three become one)
STEP_BANDS = {19: -2, 20: -2, 21: 2, 22: 2}


def _space(s):
    return SPACES[s] if isinstance(s, str) else int(s)


def colour_table(which):
    n = C.c_int()
    lib().orc_colour_table.restype = C.c_void_p
    p = lib().orc_colour_table(which, C.byref(n))
    dt = np.int32 if which in (0, 2) else np.float32
    return np.frombuffer((C.c_uint8 * (n.value * 4)).from_address(p), dtype=dt).copy()


def colour_step(a, step, interpretation):
    """One colour op (first three bands through the line function, extra bands
    rescaled / cast / re-attached as vips_colour_build does)."""
    a, h, w, b, f = _img(a)
    step = STEPS[step] if isinstance(step, str) else step
    out = np.empty((h, w, b + STEP_BANDS.get(step, 0)), STEP_IO[step][1] or a.dtype)
    if lib().orc_colour_step(step, _p(a), w, h, b, f, _space(interpretation), _p(out)):
        raise ValueError("colour_step")
    return out


def colourspace(a, space, source_space):
    a, h, w, b, f = _img(a)
    to, frm = _space(space), _space(source_space)
    out = np.empty((h, w, lib().orc_colourspace_bands(frm, to, b)), DTYPE[lib().orc_colourspace_out_format(frm, to, f)])
    if lib().orc_colourspace(_p(a), w, h, b, f, frm, to, _p(out)):
        raise ValueError("colourspace %s -> %s" % (source_space, space))
    return out
