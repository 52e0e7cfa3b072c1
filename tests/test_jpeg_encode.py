"""JPEG encode staging (the save half; csrc/jpeg_encode.cu).

vips_jpegsave_buffer (foreign/vips2jpeg.c:551-700) hands the pixels to libjpeg(-turbo) -- a third-party dependency that
is not under /root/reference -- with the library's defaults: jpeg_set_quality(Q, TRUE), JDCT_ISLOW, standard Huffman
tables, 2x2 chroma subsampling below Q 90.  The oracle is libjpeg-turbo itself (the one inside this image's Pillow, which
makes the same calls): the stream must be the same BYTES -- tables, frame header, entropy-coded segment.

CPU tests run the encoder's per-block code compiled for the host (vb200_debug_jpeg_encode); -m gpu tests the kernels.
"""
import ctypes as C
import io

import numpy as np
import pytest

PIL = pytest.importorskip("PIL.Image")

from test_jpeg import synth  # noqa: E402


def turbo_encode(a, quality, subsampling):
    b = io.BytesIO()
    PIL.fromarray(a).save(b, "JPEG", quality=quality, subsampling=subsampling)
    return b.getvalue()


def segments(d):
    """{marker: [payloads]} up to SOS, and the entropy-coded segment + EOI"""
    out, p = {}, 2
    assert d[:2] == b"\xff\xd8"
    while True:
        assert d[p] == 0xFF, p
        m = d[p + 1]
        n = (d[p + 2] << 8) | d[p + 3]
        out.setdefault(m, []).append(bytes(d[p + 4:p + 2 + n]))
        p += 2 + n
        if m == 0xDA:
            return out, bytes(d[p:])


@pytest.fixture(scope="module")
def enc():
    import libvips_b200 as vb
    L = vb.lib()
    L.vb200_debug_jpeg_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                          C.POINTER(C.c_size_t)]

    def run(a, quality, mode=0):
        a = np.ascontiguousarray(a)
        h, w = a.shape[:2]
        bands = 1 if a.ndim == 2 else a.shape[2]
        cap = w * h * 8 + 8192
        buf = (C.c_ubyte * cap)()
        n = C.c_size_t()
        vb._check(L.vb200_debug_jpeg_encode(a.ctypes.data_as(C.c_void_p), w * bands, w, h, bands, quality, mode, buf, cap, C.byref(n)))
        return bytes(buf[:n.value])
    return run


def same_stream(ours, theirs, what):
    so, eo = segments(ours)
    st, et = segments(theirs)
    assert so[0xDB] == st[0xDB], ("quantisation tables", what)
    assert so[0xC0] == st[0xC0], ("frame header", what)
    assert so[0xC4] == st[0xC4], ("huffman tables", what)
    assert so[0xDA] == st[0xDA], ("scan header", what)
    if eo != et:
        n = next(i for i in range(min(len(eo), len(et))) if eo[i] != et[i])
        raise AssertionError(("entropy-coded segment differs at byte %d of %d / %d" % (n, len(eo), len(et)), what))
    assert ours == theirs, ("whole stream", what)


@pytest.mark.parametrize("size", [(64, 64), (67, 93), (256, 200), (17, 300), (129, 31), (8, 8), (3, 5), (512, 512)], ids=lambda s: "%dx%d" % s)
def test_host_twin_writes_libjpeg_turbos_stream(enc, size):
    h, w = size
    a = synth(h, w, seed=h + 3 * w)
    for quality in (10, 50, 75, 89, 90, 100):
        # mode 0 = vips2jpeg.c's AUTO: 4:2:0 below Q 90 (PIL subsampling 2), 4:4:4 from 90 (PIL 0)
        same_stream(enc(a, quality, 0), turbo_encode(a, quality, 2 if quality < 90 else 0), (size, quality, "auto"))
    same_stream(enc(a, 75, 2), turbo_encode(a, 75, 0), (size, 75, "subsample off"))
    same_stream(enc(a, 95, 1), turbo_encode(a, 95, 2), (size, 95, "subsample on"))
    g = synth(h, w, seed=w, grey=True)
    same_stream(enc(g, 75, 0), turbo_encode(g, 75, 0), (size, 75, "grey"))     # vips2jpeg.c:678-684: one band is always 1 x 1


def test_extremes(enc):
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:96, 0:128]
    check = np.repeat((((yy + xx) % 2) * 255).astype(np.uint8)[..., None], 3, -1)
    flat = np.full((40, 56, 3), 255, np.uint8)
    for a in (noise, check, flat, 255 - flat):
        for q in (1, 30, 100):
            same_stream(enc(a, q, 0), turbo_encode(a, q, 2 if q < 90 else 0), (a.shape, q))


def test_round_trip_through_the_device_decoder_twin(enc):
    """what the encoder writes, the decoder (csrc/jpeg.cu) reads back as libjpeg-turbo would"""
    import libvips_b200 as vb
    a = synth(120, 200, seed=8)
    d = enc(a, 85, 0)
    got = vb.jpeg_decode_host_twin(d, 1)
    want = np.asarray(PIL.open(io.BytesIO(d)))
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_batch_writes_libjpeg_turbos_streams(enc):
    import libvips_b200 as vb
    vb.init(0)
    for (h, w) in ((64, 64), (67, 93), (256, 200), (17, 300), (8, 8), (512, 512)):
        frames = np.stack([synth(h, w, seed=i + h) for i in range(3)])
        for quality, mode, sub in ((75, "auto", 2), (92, "auto", 0), (10, "auto", 2), (100, "on", 2), (60, "off", 0)):
            got = vb.jpegsave_batch(frames, quality, mode)
            for i in range(3):
                same_stream(got[i], turbo_encode(frames[i], quality, sub), ((h, w), quality, mode, i))
        grey = frames[..., 0].copy()
        got = vb.jpegsave_batch(grey, 80)
        for i in range(3):
            same_stream(got[i], turbo_encode(grey[i], 80, 0), ((h, w), "grey", i))
    rng = np.random.default_rng(4)
    noise = rng.integers(0, 256, (2, 96, 128, 3), dtype=np.uint8)
    for q in (1, 100):
        got = vb.jpegsave_batch(noise, q)
        for i in range(2):
            same_stream(got[i], turbo_encode(noise[i], q, 2 if q < 90 else 0), ("noise", q, i))
    with pytest.raises(vb.Error, match="stride"):
        vb.jpegsave_batch(noise, 100, stride=4096)


@pytest.mark.gpu
def test_gpu_jpeg_in_jpeg_out(enc):
    """the thumbnail server's whole loop on the device: JPEG streams -> shrink-on-load -> thumbnail -> JPEG streams"""
    import torch
    import libvips_b200 as vb
    from oracle import pyoracle
    from test_jpeg import encode, turbo_decode
    vb.init(0)
    h, w, target = 1024, 1536, 256
    streams = [encode(synth(h, w, seed=i), 88, 2) for i in range(3)]
    shrink = vb.thumbnail_jpegshrink(w, h, target)
    dw, dh, bands = vb.jpeg_geometry(streams, shrink)
    plan = vb.ThumbnailPlan(dw, dh, bands, target)
    out = torch.empty((3, plan.out_height, plan.out_width, bands), dtype=torch.uint8, device="cuda")
    plan.run_jpeg(streams, shrink, out_ptr=out.data_ptr())
    torch.cuda.synchronize()
    got = vb.jpegsave_batch(None, 75, in_ptr=out.data_ptr(), shape=tuple(out.shape))
    for i in range(3):
        thumb = pyoracle.thumbnail_image(turbo_decode(streams[i], shrink), target)
        same_stream(got[i], turbo_encode(thumb, 75, 2), i)

