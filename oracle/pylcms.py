"""pylcms.py -- TEST INFRASTRUCTURE ONLY: the ICC rows of the oracle.

vips_icc_import / vips_icc_export / vips_icc_transform do their arithmetic inside lcms2
(colour/icc_transform.c:459-463 cmsCreateTransform, :931/:1094/:1114/:1219 cmsDoTransform),
which is not under /root/reference.  The reference pins no lcms2 version (meson.build:444-447).
This module binds the liblcms2 2.18 that ships inside the Pillow wheel of this image with
ctypes and makes EXACTLY the calls the reference makes -- the same profiles
(cmsCreateLab4Profile(D65 white from 6504 K), cmsCreateXYZProfile), the same pixel formats
(vips_icc_make_lcms_format), cmsFLAGS_NOCACHE, the same decode_lab / decode_xyz / encode_xyz
around them.  Parity of the CUDA ICC path is pinned to it within a stated tolerance (lcms2's
own 8-bit-input transforms are table-interpolated, so bit-exactness with a colorimetric
evaluator is not a meaningful bar; the reference's tests use dE < 6 / |diff| < 3).

Product code never imports this."""
import ctypes as C
import glob
import os

import numpy as np

PT_GRAY, PT_RGB, PT_CMYK, PT_XYZ, PT_Lab = 3, 4, 6, 9, 10
NOCACHE = 0x0040
INTENTS = {"perceptual": 0, "relative": 1, "saturation": 2, "absolute": 3}
_SIG = {b"GRAY": (1, PT_GRAY), b"RGB ": (3, PT_RGB), b"CMYK": (4, PT_CMYK), b"Lab ": (3, PT_Lab), b"XYZ ": (3, PT_XYZ)}

_L = None


def lib():
    global _L
    if _L is None:
        import PIL
        hits = glob.glob(os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs", "liblcms2*"))
        if not hits:
            raise OSError("no liblcms2 next to Pillow")
        L = C.CDLL(hits[0])
        L.cmsOpenProfileFromMem.restype = C.c_void_p
        L.cmsOpenProfileFromMem.argtypes = [C.c_char_p, C.c_uint]
        L.cmsCreateLab4Profile.restype = C.c_void_p
        L.cmsCreateLab4Profile.argtypes = [C.c_void_p]
        L.cmsCreateXYZProfile.restype = C.c_void_p
        L.cmsWhitePointFromTemp.argtypes = [C.c_void_p, C.c_double]
        L.cmsCreateTransform.restype = C.c_void_p
        L.cmsCreateTransform.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_uint, C.c_uint]
        L.cmsDoTransform.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
        L.cmsDeleteTransform.argtypes = [C.c_void_p]
        L.cmsCloseProfile.argtypes = [C.c_void_p]
        L.cmsGetEncodedCMMversion.restype = C.c_int
        _L = L
    return _L


def available():
    try:
        return lib().cmsGetEncodedCMMversion()
    except OSError:
        return 0


def _fmt(pixel_type, bands, nbytes):
    """vips_icc_make_lcms_format, icc_transform.c:262-269"""
    return ((1 if nbytes == 4 else 0) << 22) | (pixel_type << 16) | (bands << 3) | (nbytes & 7)


def _device(profile):
    cs = profile[16:20]
    return _SIG[cs]


def _pcs_profile(pcs):
    L = lib()
    if pcs == "lab":
        white = (C.c_double * 3)()
        L.cmsWhitePointFromTemp(white, 6504.0)          # icc_transform.c:823-827
        return L.cmsCreateLab4Profile(white), _SIG[b"Lab "]
    return L.cmsCreateXYZProfile(), _SIG[b"XYZ "]


def _run(pin, fin, pout, fout, intent, src, dst, n):
    L = lib()
    t = L.cmsCreateTransform(pin, fin, pout, fout, INTENTS[intent], NOCACHE)
    if not t:
        raise ValueError("cmsCreateTransform failed")
    L.cmsDoTransform(t, src.ctypes.data, dst.ctypes.data, n)
    L.cmsDeleteTransform(t)


def _nbytes(dt):
    return {np.dtype(np.uint8): 1, np.dtype(np.uint16): 2, np.dtype(np.float32): 4}[np.dtype(dt)]


def icc_import(a, profile, intent="relative", pcs="lab"):
    """vips_icc_import: device (u8 / u16 / f32) -> PCS float, icc_transform.c:813-945."""
    L = lib()
    a = np.ascontiguousarray(a)
    bands, pt = _device(profile)
    assert a.shape[-1] == bands
    n = a.size // bands
    hin = L.cmsOpenProfileFromMem(profile, len(profile))
    hout, (_, ptp) = _pcs_profile(pcs)
    enc = np.zeros((n, 3), np.uint16)
    _run(hin, _fmt(pt, bands, _nbytes(a.dtype)), hout, _fmt(ptp, 3, 2), intent, a, enc, n)
    L.cmsCloseProfile(hin)
    L.cmsCloseProfile(hout)
    out = np.empty((n, 3), np.float32)
    f = enc.astype(np.float64)
    if pcs == "lab":                                      # decode_lab :856-872
        out[:, 0] = f[:, 0] / 655.35
        out[:, 1] = f[:, 1] / 257.0 - 128.0
        out[:, 2] = f[:, 2] / 257.0 - 128.0
    else:                                                 # decode_xyz :879-909 (float arithmetic)
        s = np.float32(100.0)
        x = (f / 32768.0).astype(np.float32) * s
        X, Y, Z = x[:, 0], x[:, 1], x[:, 2]
        f32 = np.float32
        out[:, 0] = f32(0.955513) * X + f32(-0.023073) * Y + f32(0.063309) * Z
        out[:, 1] = f32(-0.028325) * X + f32(1.009942) * Y + f32(0.021055) * Z
        out[:, 2] = f32(0.012329) * X + f32(-0.020536) * Y + f32(1.330714) * Z
    return out.reshape(a.shape[:-1] + (3,))


def icc_export(p, profile, intent="relative", depth=8, pcs="lab"):
    """vips_icc_export: PCS float -> device u8 / u16, icc_transform.c:995-1117."""
    L = lib()
    p = np.ascontiguousarray(p, np.float32)
    bands, pt = _device(profile)
    n = p.size // 3
    hin, (_, ptp) = _pcs_profile(pcs)
    hout = L.cmsOpenProfileFromMem(profile, len(profile))
    if pcs == "xyz":                                      # encode_xyz :1050-1076
        f32 = np.float32
        x = (p.reshape(-1, 3) / f32(100.0)).astype(np.float32)
        X, Y, Z = x[:, 0], x[:, 1], x[:, 2]
        q = np.empty((n, 3), np.float32)
        q[:, 0] = f32(1.047886) * X + f32(0.022919) * Y + f32(-0.050216) * Z
        q[:, 1] = f32(0.029582) * X + f32(0.990484) * Y + f32(-0.017079) * Z
        q[:, 2] = f32(-0.009252) * X + f32(0.015073) * Y + f32(0.751678) * Z
        src = q
    else:
        src = p.reshape(-1, 3)
    dt = np.uint8 if depth == 8 else np.uint16
    out = np.zeros((n, bands), dt)
    _run(hin, _fmt(ptp, 3, 4), hout, _fmt(pt, bands, depth // 8), intent, np.ascontiguousarray(src), out, n)
    L.cmsCloseProfile(hin)
    L.cmsCloseProfile(hout)
    return out.reshape(p.shape[:-1] + (bands,))


def icc_transform(a, in_profile, out_profile, intent="relative", depth=8):
    """vips_icc_transform: device -> device with one lcms2 transform, icc_transform.c:1166-1220."""
    L = lib()
    a = np.ascontiguousarray(a)
    bi, pti = _device(in_profile)
    bo, pto = _device(out_profile)
    n = a.size // bi
    hin = L.cmsOpenProfileFromMem(in_profile, len(in_profile))
    hout = L.cmsOpenProfileFromMem(out_profile, len(out_profile))
    dt = np.uint8 if depth == 8 else np.uint16
    out = np.zeros((n, bo), dt)
    _run(hin, _fmt(pti, bi, _nbytes(a.dtype)), hout, _fmt(pto, bo, depth // 8), intent, a, out, n)
    L.cmsCloseProfile(hin)
    L.cmsCloseProfile(hout)
    return out.reshape(a.shape[:-1] + (bo,))
