"""ICC profiles built from scratch for the ICC tests (nothing copied from the reference tree):
an sRGB-like v4 matrix / parametric-TRC profile, a gamma and a table TRC variant, a grey TRC
profile and a synthetic 4-ink lut16 profile with a Lab PCS (ICC.1:2001-04 lut16Type)."""
import struct

import numpy as np


def _s15(v):
    return struct.pack(">i", int(round(v * 65536.0)))


def _xyz_tag(x, y, z):
    return b"XYZ " + b"\0" * 4 + _s15(x) + _s15(y) + _s15(z)


def _para(ftype, params):
    return b"para" + b"\0" * 4 + struct.pack(">HH", ftype, 0) + b"".join(_s15(p) for p in params)


def _curv(values):
    """values: [] identity, [gamma] u8.8 gamma, or a table of floats 0..1"""
    if len(values) == 1:
        return b"curv" + b"\0" * 4 + struct.pack(">I", 1) + struct.pack(">H", int(round(values[0] * 256)))
    body = b"".join(struct.pack(">H", int(round(v * 65535))) for v in values)
    return b"curv" + b"\0" * 4 + struct.pack(">I", len(values)) + body


def _text(sig, s):
    return b"text" + b"\0" * 4 + s.encode() + b"\0"


def _profile(version, cls, cs, pcs, tags):
    """tags: list of (sig, bytes); identical payloads are not shared (keeps this short)"""
    n = len(tags)
    off = 128 + 4 + 12 * n
    table, body = b"", b""
    for sig, data in tags:
        pad = (-len(data)) % 4
        table += sig.encode() + struct.pack(">II", off + len(body), len(data))
        body += data + b"\0" * pad
    size = off + len(body)
    hdr = struct.pack(">I4sI4s4s4s", size, b"vb2h", version, cls.encode(), cs.encode(), pcs.encode())
    hdr += struct.pack(">6H", 2024, 1, 1, 0, 0, 0) + b"acsp" + b"APPL" + struct.pack(">I", 0)
    hdr += b"\0" * 4 + b"\0" * 4 + b"\0" * 8 + struct.pack(">I", 1)          # manufacturer, model, attributes, intent
    hdr += _s15(0.9642) + _s15(1.0) + _s15(0.8249) + b"vb2h" + b"\0" * 16     # illuminant D50, creator, id
    hdr += b"\0" * (128 - len(hdr))
    return hdr + struct.pack(">I", n) + table + body


# Bradford-adapted sRGB primaries (D50), as in every v4 sRGB profile
_SRGB_COL = [(0.43607, 0.22249, 0.01392), (0.38515, 0.71687, 0.09708), (0.14307, 0.06061, 0.71410)]
_SRGB_PARA = (3, [2.4, 1 / 1.055, 0.055 / 1.055, 1 / 12.92, 0.04045])


def rgb_profile(trc="srgb"):
    if trc == "srgb":
        curve = _para(*_SRGB_PARA)
    elif trc == "gamma":
        curve = _curv([2.2])
    elif trc == "para4":
        curve = _para(4, [2.2, 0.95, 0.05, 0.223, 0.1, 0.01, 0.002])  # continuous at X = d to 1e-4
    else:
        x = np.linspace(0, 1, 1024)
        curve = _curv(list(x ** 1.8))
    tags = [("desc", _text("desc", "vb200 test rgb")), ("cprt", _text("cprt", "none")), ("wtpt", _xyz_tag(0.9642, 1.0, 0.8249))]
    for name, col in zip("rgb", _SRGB_COL):
        tags.append((name + "XYZ", _xyz_tag(*col)))
    for name in "rgb":
        tags.append((name + "TRC", curve))
    return _profile(0x04200000, "mntr", "RGB ", "XYZ ", tags)


def grey_profile():
    tags = [("desc", _text("desc", "vb200 test grey")), ("cprt", _text("cprt", "none")), ("wtpt", _xyz_tag(0.9642, 1.0, 0.8249)),
            ("kTRC", _para(*_SRGB_PARA))]
    return _profile(0x04200000, "mntr", "GRAY", "XYZ ", tags)


def _lab_to_v2(lab):
    """ICC v2 lut16 Lab encoding as 0..65535"""
    out = np.empty_like(lab)
    out[..., 0] = lab[..., 0] / 100.0 * 65280.0
    out[..., 1:] = (lab[..., 1:] + 128.0) * 256.0
    return np.clip(np.rint(out), 0, 65535)


def _xyz_to_lab(xyz):
    t = xyz / np.array([0.9642, 1.0, 0.8249])
    f = np.where(t > 216 / 24389, np.cbrt(t), (841 / 108) * t + 16 / 116)
    return np.stack([116 * f[..., 1] - 16, 500 * (f[..., 0] - f[..., 1]), 200 * (f[..., 1] - f[..., 2])], -1)


def _lab_to_xyz(lab):
    fy = (lab[..., 0] + 16) / 116
    fx, fz = fy + lab[..., 1] / 500, fy - lab[..., 2] / 200
    f = np.stack([fx, fy, fz], -1)
    t = np.where(f > 24 / 116, f ** 3, (108 / 841) * (f - 16 / 116))
    return t * np.array([0.9642, 1.0, 0.8249])


def _mft2(in_ch, out_ch, grid, in_tables, clut, out_tables):
    ident = [1, 0, 0, 0, 1, 0, 0, 0, 1]
    b = b"mft2" + b"\0" * 4 + bytes([in_ch, out_ch, grid, 0]) + b"".join(_s15(v) for v in ident)
    b += struct.pack(">HH", in_tables.shape[1], out_tables.shape[1])
    for arr in (in_tables, clut, out_tables):
        b += np.ascontiguousarray(arr, dtype=">u2").tobytes()
    return b


_M = np.array(_SRGB_COL).T  # linear rgb -> XYZ D50


def ink_profile(grid_a2b=9, grid_b2a=17):
    """A smooth 4-ink device: rgb = (1 - cmy) * (1 - 0.9 k), gamma 2, through the sRGB primaries."""
    ramp = np.linspace(0, 65535, 256)[None, :]
    g = np.linspace(0, 1, grid_a2b)
    C, Mg, Y, K = np.meshgrid(g, g, g, g, indexing="ij")                     # first channel slowest
    rgb = (1 - np.stack([C, Mg, Y], -1)) * (1 - 0.9 * K[..., None])
    lab = _xyz_to_lab((rgb ** 2.0) @ _M.T)
    a2b = _mft2(4, 3, grid_a2b, np.repeat(ramp, 4, 0), _lab_to_v2(lab).reshape(-1, 3), np.repeat(ramp, 3, 0))
    # B2A: Lab (v2 encoded grid) -> inks with k = 0, out-of-gamut clipped
    e = np.linspace(0, 65535, grid_b2a)
    Le, ae, be = np.meshgrid(e, e, e, indexing="ij")
    lab = np.stack([Le / 65280.0 * 100.0, ae / 256.0 - 128.0, be / 256.0 - 128.0], -1)
    lin = np.clip(_lab_to_xyz(lab) @ np.linalg.inv(_M).T, 0, 1)
    cmy = 1 - np.sqrt(lin)
    inks = np.concatenate([cmy, np.zeros_like(cmy[..., :1])], -1)
    b2a = _mft2(3, 4, grid_b2a, np.repeat(ramp, 3, 0), np.rint(inks * 65535).reshape(-1, 4), np.repeat(ramp, 4, 0))
    tags = [("desc", _text("desc", "vb200 test inks")), ("cprt", _text("cprt", "none")), ("wtpt", _xyz_tag(0.9642, 1.0, 0.8249)),
            ("A2B0", a2b), ("A2B1", a2b), ("B2A0", b2a), ("B2A1", b2a)]
    return _profile(0x02200000, "prtr", "CMYK", "Lab ", tags)


def _mab(sig, in_ch, out_ch, b_curves, matrix=None, m_curves=None, clut=None, a_curves=None):
    """lutAtoBType ('mAB ') / lutBtoAType ('mBA '), ICC.1:2010 10.10 / 10.11.  clut = (grid tuple, array[..., out] 0..1)."""
    def pad(x):
        return x + b"\0" * ((-len(x)) % 4)
    parts, off = {}, 32
    body = b""
    def put(name, data):
        nonlocal body, off
        parts[name] = off
        body += pad(data)
        off = 32 + len(body)
    put("b", b"".join(pad(c) for c in b_curves))
    if matrix is not None:
        put("mat", b"".join(_s15(v) for v in matrix))
    if m_curves is not None:
        put("m", b"".join(pad(c) for c in m_curves))
    if clut is not None:
        grid, arr = clut
        g = bytes(list(grid) + [0] * (16 - len(grid)))
        put("clut", g + bytes([2, 0, 0, 0]) + np.ascontiguousarray(np.rint(np.clip(arr, 0, 1) * 65535), dtype=">u2").tobytes())
    if a_curves is not None:
        put("a", b"".join(pad(c) for c in a_curves))
    hdr = sig + b"\0" * 4 + bytes([in_ch, out_ch, 0, 0])
    hdr += struct.pack(">5I", parts["b"], parts.get("mat", 0), parts.get("m", 0), parts.get("clut", 0), parts.get("a", 0))
    return hdr + body


def lut_v4_rgb_profile(pcs="XYZ "):
    """A v4 RGB profile whose transforms are lutAtoB / lutBtoA tags: gamma-2 A curves, a gently bent CLUT on a
    non-uniform grid, identity M curves, the sRGB matrix (in the tag's encoded PCS units) and identity B curves."""
    ident = _curv([])
    grid = (9, 7, 8)
    axes = [np.linspace(0, 1, g) for g in grid]
    R, G, B = np.meshgrid(*axes, indexing="ij")
    lin = np.stack([R + 0.04 * np.sin(np.pi * R) * G, G + 0.03 * np.sin(np.pi * G) * B, B - 0.03 * np.sin(np.pi * B) * R], -1)
    if pcs == "XYZ ":
        enc = 32768.0 / 65535.0
        mat = list((_M * enc).reshape(-1)) + [0.0, 0.0, 0.0]
        a2b = _mab(b"mAB ", 3, 3, [ident] * 3, matrix=mat, m_curves=[ident] * 3, clut=(grid, lin), a_curves=[_para(0, [2.0])] * 3)
        inv = list((np.linalg.inv(_M) / enc).reshape(-1)) + [0.0, 0.0, 0.0]
        g2 = (11, 11, 11)
        ax = np.linspace(0, 1, 11)
        X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
        ident_clut = np.stack([X, Y, Z], -1)
        b2a = _mab(b"mBA ", 3, 3, [ident] * 3, matrix=inv, m_curves=[ident] * 3, clut=(g2, ident_clut), a_curves=[_para(0, [0.5])] * 3)
    else:
        # Lab PCS: the CLUT holds encoded Lab directly (no matrix), B curves identity
        lab = _xyz_to_lab(np.clip(lin, 0, 1) @ _M.T)
        enc_lab = np.stack([lab[..., 0] / 100.0, (lab[..., 1] + 128.0) / 255.0, (lab[..., 2] + 128.0) / 255.0], -1)
        a2b = _mab(b"mAB ", 3, 3, [ident] * 3, clut=(grid, enc_lab), a_curves=[_para(0, [2.0])] * 3)
        g2 = (13, 13, 13)
        ax = np.linspace(0, 1, 13)
        Le, ae, be = np.meshgrid(ax, ax, ax, indexing="ij")
        labg = np.stack([Le * 100.0, ae * 255.0 - 128.0, be * 255.0 - 128.0], -1)
        linb = np.clip(_lab_to_xyz(labg) @ np.linalg.inv(_M).T, 0, 1)
        b2a = _mab(b"mBA ", 3, 3, [ident] * 3, clut=(g2, linb), a_curves=[_para(0, [0.5])] * 3)
    tags = [("desc", _text("desc", "vb200 test lut v4")), ("cprt", _text("cprt", "none")), ("wtpt", _xyz_tag(0.9642, 1.0, 0.8249)),
            ("A2B0", a2b), ("A2B1", a2b), ("B2A0", b2a), ("B2A1", b2a)]
    return _profile(0x04200000, "mntr", "RGB ", pcs, tags)
