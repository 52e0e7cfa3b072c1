"""ctypes loader for oracle/_ref/libvipsref.so: the reference's OWN source files
(libvips resample/*.c*, conversion/{pre,unpre}multiply.c, ...) compiled in place
under the GLib-free shim (oracle/ref_shim).  TEST INFRASTRUCTURE ONLY.

Pipelines are built lazily exactly like libvips does (build() per op, then a
sink pulls tiles through the generate() callbacks), so the rects each generate
function sees are the ones the real library would produce.
"""
import ctypes as C
import os

import numpy as np

from . import pyoracle

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libvipsref.so")
_LIB = None


def available():
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(PATH)
        for name in ("ref_image_new_from_memory", "ref_shrinkv", "ref_shrinkh", "ref_reducev", "ref_reduceh",
                     "ref_resize", "ref_premultiply", "ref_unpremultiply", "ref_colour_op", "ref_colourspace",
                     "ref_conv", "ref_convsep", "ref_gaussblur", "ref_sharpen", "ref_gaussmat", "ref_affine",
                     "ref_cast", "ref_colourspace_build"):
            if hasattr(L, name):
                getattr(L, name).restype = C.c_void_p
        L.ref_error.restype = C.c_char_p
        L.ref_image_new_from_memory.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_shrinkv.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_shrinkh.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_reducev.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double]
        L.ref_reduceh.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double]
        L.ref_resize.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double]
        L.ref_premultiply.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.ref_unpremultiply.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.ref_image_write_to_memory.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        for f in ("ref_image_width", "ref_image_height", "ref_image_bands", "ref_image_format", "ref_image_dhint"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_reduce_make_mask.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.vips_reduce_get_points.argtypes = [C.c_int, C.c_double]
        if hasattr(L, "ref_colourspace_build"):
            L.ref_colourspace_build.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.ref_cast.argtypes = [C.c_void_p, C.c_int, C.c_int]
        if hasattr(L, "ref_thumbnail_calculate_shrink"):
            L.ref_thumbnail_calculate_shrink.restype = None
            L.ref_thumbnail_calculate_shrink.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_double)] * 2
            L.ref_thumbnail_find_jpegshrink.argtypes = [C.c_int] * 7
        _LIB = L
    return _LIB


class RefImage:
    """A lazy reference image (VipsImage* inside the shim)."""

    def __init__(self, handle, keep=()):
        if not handle:
            raise ValueError("reference op failed: %s" % lib().ref_error().decode())
        self.h = handle
        self.keep = keep  # keep source arrays alive

    @staticmethod
    def from_array(a, interpretation=None):
        a = np.ascontiguousarray(a)
        if a.ndim == 2:
            a = a[:, :, None]
        if interpretation is None:
            interpretation = 1 if a.shape[2] < 3 else 22
        h = lib().ref_image_new_from_memory(a.ctypes.data, a.shape[1], a.shape[0], a.shape[2],
                                            pyoracle.FMT[a.dtype], interpretation)
        return RefImage(h, (a,))

    def _op(self, fn, *args):
        return RefImage(fn(self.h, *args), (self,))

    def shrinkv(self, f, ceil=False):
        return self._op(lib().ref_shrinkv, f, int(ceil))

    def shrinkh(self, f, ceil=False):
        return self._op(lib().ref_shrinkh, f, int(ceil))

    def reducev(self, f, kernel="lanczos3", gap=0.0):
        return self._op(lib().ref_reducev, float(f), pyoracle._k(kernel), float(gap))

    def reduceh(self, f, kernel="lanczos3", gap=0.0):
        return self._op(lib().ref_reduceh, float(f), pyoracle._k(kernel), float(gap))

    def resize(self, scale, vscale=None, kernel="lanczos3", gap=2.0):
        return self._op(lib().ref_resize, float(scale), float(scale if vscale is None else vscale),
                        pyoracle._k(kernel), float(gap))

    def premultiply(self, max_alpha=0.0, uchar=False):
        return self._op(lib().ref_premultiply, float(max_alpha), int(uchar))

    def unpremultiply(self, max_alpha=0.0, uchar=False):
        return self._op(lib().ref_unpremultiply, float(max_alpha), int(uchar))

    @property
    def shape(self):
        L = lib()
        return (L.ref_image_height(self.h), L.ref_image_width(self.h), L.ref_image_bands(self.h))

    @property
    def dhint(self):
        return lib().ref_image_dhint(self.h)

    def colourspace(self, space, source_space):
        """vips_colourspace_build over colourspace.c's route table, every step a real VipsColour object (colour.c build,
        vips_colour_gen, the converter's line function); alpha through cast.c / linear.c"""
        return self._op(lib().ref_colourspace_build, pyoracle._space(space), pyoracle._space(source_space))

    def cast(self, dtype, shift=False):
        return self._op(lib().ref_cast, pyoracle.FMT[np.dtype(dtype)], int(shift))

    def numpy(self, tile=(0, 0)):
        """The sink: pull the image through generate() tile by tile."""
        out = np.empty(self.shape, pyoracle.DTYPE[lib().ref_image_format(self.h)])
        if lib().ref_image_write_to_memory(self.h, out.ctypes.data, tile[0], tile[1]):
            raise ValueError("reference evaluation failed: %s" % lib().ref_error().decode())
        return out


def reduce_make_mask(kernel, n_point, shrink, x):
    c = np.zeros(n_point, np.float64)
    lib().ref_reduce_make_mask(c.ctypes.data, pyoracle._k(kernel), n_point, shrink, x)
    return c


def reduce_get_points(kernel, shrink):
    return lib().vips_reduce_get_points(pyoracle._k(kernel), shrink)


def thumbnail_image(a, width, height=None, size="both", tile=(0, 0)):
    """vips_thumbnail_build's pixel chain for an 8-bit sRGB / B_W image
    (thumbnail.c:827-902): [premultiply uchar] -> resize -> [unpremultiply uchar]."""
    h, w, b = a.shape
    hs, vs = thumbnail_calculate_shrink(w, h, width, height, size)      # the reference's own arithmetic
    im = RefImage.from_array(a)
    premul = (b == 2 or b >= 4) and hs != 1.0 and vs != 1.0
    if premul:
        im = im.premultiply(uchar=True)
    im = im.resize(1.0 / hs, 1.0 / vs)
    if premul:
        im = im.unpremultiply(uchar=True)
    return im.numpy(tile)


# ------------------------------------------------------------------ colour
def colour_line(step, a):
    """Run the reference's own *_line function over an (n, 3) array."""
    step = pyoracle.STEPS[step] if isinstance(step, str) else step
    a = np.ascontiguousarray(a).reshape(-1, 3)
    assert a.dtype == pyoracle.STEP_IO[step][0], (a.dtype, step)
    out = np.empty((a.shape[0], 3 + pyoracle.STEP_BANDS.get(step, 0)), pyoracle.STEP_IO[step][1])
    L = lib()
    L.ref_colour_line.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    if L.ref_colour_line(step, a.ctypes.data, out.ctypes.data, a.shape[0]):
        raise ValueError("ref_colour_line")
    return out


def colour_table(which):
    L = lib()
    L.ref_colour_table.restype = C.c_void_p
    n = C.c_int()
    p = L.ref_colour_table(which, C.byref(n))
    dt = np.int32 if which in (0, 2) else np.float32
    return np.frombuffer((C.c_uint8 * (n.value * 4)).from_address(p), dtype=dt).copy()


# ------------------------------------------------------------------ vips_thumbnail's size arithmetic
def thumbnail_calculate_shrink(w, h, width, height=None, size="both", crop=0):
    """vips_thumbnail_calculate_shrink (thumbnail.c:412-466), the file-static function itself (ref_shim/ref_thumbnail.c)"""
    hs, vs = C.c_double(), C.c_double()
    lib().ref_thumbnail_calculate_shrink(w, h, width, width if height is None else height, pyoracle.SIZES[size], crop,
                                         C.byref(hs), C.byref(vs))
    return hs.value, vs.value


def thumbnail_find_jpegshrink(w, h, width, height=None, size="both", crop=0, linear=False):
    """vips_thumbnail_find_jpegshrink (thumbnail.c:490-517)"""
    return int(lib().ref_thumbnail_find_jpegshrink(w, h, width, width if height is None else height, pyoracle.SIZES[size], crop,
                                                   int(linear)))
