"""JPEG decode staging with shrink-on-load (csrc/jpeg.cu).

The reference loads JPEGs through libjpeg (foreign/jpeg2vips.c), a third-party dependency that is not under
/root/reference.  The oracle for this row is therefore libjpeg-turbo ITSELF, as shipped inside this image's Pillow:
PIL decodes with the library's defaults (JDCT_ISLOW, fancy upsampling -- what jpeg2vips.c uses) and its draft mode
sets scale_denom exactly as jpeg2vips.c:537-538 does.  Bit for bit, no tolerance.

CPU tests run the decoder's per-block code compiled for the host (vb200_debug_jpeg_decode); the -m gpu tests run the
kernels through the C ABI, and the thumbnail feed against the oracle thumbnail of libjpeg-turbo's decode.
"""
import io

import numpy as np
import pytest

PIL = pytest.importorskip("PIL.Image")
from PIL import ImageFile as _ImageFile  # noqa: E402

_ImageFile.MAXBLOCK = 1 << 24   # Pillow's encoder buffer: progressive / optimised saves at Q 100 outgrow the default


def synth(h, w, seed=0, grey=False):
    """smooth structure + noise: compresses like a photograph, exercises every coefficient position"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(xx / 37.0 + yy / 91.0), 128 + 90 * np.cos(xx / 53.0 - yy / 29.0), (xx * 3 + yy * 5) % 256], -1)
    base = base + rng.normal(0, 12, base.shape)
    a = np.clip(base, 0, 255).astype(np.uint8)
    return a[..., 0] if grey else a


def encode(a, quality=85, subsampling=2, **kw):
    b = io.BytesIO()
    PIL.fromarray(a).save(b, "JPEG", quality=quality, subsampling=subsampling, **kw)
    return b.getvalue()


def turbo_decode(data, shrink):
    """libjpeg-turbo with scale_num / scale_denom = 1 / shrink, cropped as jpeg2vips.c:639-640 crops it"""
    im = PIL.open(io.BytesIO(data))
    w, h = im.size
    if shrink > 1:
        im.draft(im.mode, (max(1, w // shrink), max(1, h // shrink)))
        assert im.size == ((w + shrink - 1) // shrink, (h + shrink - 1) // shrink), "draft() picked another scale"
    a = np.asarray(im)[: h // shrink, : w // shrink]
    return a[..., None] if a.ndim == 2 else a


@pytest.fixture(scope="module")
def vbl():
    import libvips_b200 as vb
    vb.lib()
    return vb


CASES = [(64, 64), (67, 93), (256, 200), (17, 300), (129, 31)]


CASES += [(8, 3), (3, 40), (16, 16), (33, 5)]


@pytest.mark.parametrize("size", CASES, ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("sub", [2, 1, 0], ids=["420", "422", "444"])
def test_host_twin_matches_libjpeg_turbo(vbl, size, sub):
    """every sampling x shrink: where libjpeg's upsampler is the identity (one kernel per MCU) and where it is not
    (4:2:0 at full size: h2v2 fancy; 4:2:2: h2v1 fancy; plain replication at 1/8 and for components <= 2 samples wide)"""
    h, w = size
    a = synth(h, w, seed=h * 7 + w)
    for quality in (30, 85, 100):
        d = encode(a, quality, sub)
        for shrink in (8, 4, 2, 1):
            if min(h, w) // shrink < 1:
                continue
            got = vbl.jpeg_decode_host_twin(d, shrink)
            want = turbo_decode(d, shrink)
            assert got.shape == want.shape and np.array_equal(got, want), (size, sub, quality, shrink)


@pytest.mark.parametrize("size", [(64, 64), (67, 93), (256, 200), (17, 300), (8, 3), (33, 5)], ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("sub", [2, 1, 0], ids=["420", "422", "444"])
def test_progressive_host_twin_matches_libjpeg_turbo(vbl, size, sub):
    """progressive frames (T.81 G: DC scans, AC bands per component, successive approximation with refinement scans) end in
    the same coefficients, so the same pixels, as libjpeg-turbo decodes -- at every shrink"""
    h, w = size
    a = synth(h, w, seed=h + 11 * w)
    for quality in (30, 85, 100):
        for kw in ({}, {"restart_marker_rows": 1}):
            d = encode(a, quality, sub, progressive=True, **kw)
            for shrink in (8, 4, 2, 1):
                if min(h, w) // shrink < 1:
                    continue
                got = vbl.jpeg_decode_host_twin(d, shrink)
                want = turbo_decode(d, shrink)
                assert got.shape == want.shape and np.array_equal(got, want), (size, sub, quality, kw, shrink)
    g = encode(synth(h, w, seed=5, grey=True), 80, progressive=True)
    assert np.array_equal(vbl.jpeg_decode_host_twin(g, 1), turbo_decode(g, 1))


def test_greyscale_restart_markers_and_optimised_tables(vbl):
    g = synth(150, 203, seed=3, grey=True)
    d = encode(g, 90)
    for shrink in (1, 2, 4, 8):
        assert np.array_equal(vbl.jpeg_decode_host_twin(d, shrink), turbo_decode(d, shrink))
    a = synth(200, 312, seed=4)
    for kw in ({"restart_marker_rows": 1}, {"restart_marker_blocks": 3}, {"optimize": True}, {"optimize": True, "restart_marker_rows": 2}):
        for sub in (2, 0):
            d = encode(a, 80, sub, **kw)
            for shrink in (2, 4, 8):
                assert np.array_equal(vbl.jpeg_decode_host_twin(d, shrink), turbo_decode(d, shrink)), (kw, sub, shrink)


def test_extreme_coefficients(vbl):
    """noise at quality 100 and saturated checkerboards: the range-limit table's clamps and the widest Huffman codes"""
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:96, 0:128]
    check = np.repeat((((yy + xx) % 2) * 255).astype(np.uint8)[..., None], 3, -1)
    for a in (noise, check):
        for q in (100, 5):
            for sub in (2, 0):
                d = encode(a, q, sub)
                for shrink in (1, 2, 4, 8):
                    assert np.array_equal(vbl.jpeg_decode_host_twin(d, shrink), turbo_decode(d, shrink)), (q, sub, shrink)


def test_declined_streams(vbl):
    a = synth(64, 64)
    ycck = io.BytesIO()
    PIL.fromarray(np.dstack([a, a[..., 0]]), "CMYK").save(ycck, "JPEG")
    with pytest.raises(vbl.Error, match="component"):
        vbl.jpeg_decode_host_twin(ycck.getvalue(), 1)               # 4 components
    with pytest.raises(vbl.Error, match="shrink"):
        vbl.jpeg_decode_host_twin(encode(a), 3)
    with pytest.raises(vbl.Error, match="JPEG"):
        vbl.jpeg_decode_host_twin(b"\x89PNG\r\n\x1a\n" + bytes(64), 1)
    d = encode(a, subsampling=0)
    for cut in (2, 20, 200, len(d) // 2):
        # truncated headers are errors; a truncated scan decodes (zeros fed past the end, as jdhuff.c does) or reports a bad code
        try:
            vbl.jpeg_decode_host_twin(d[:cut], 1)
        except vbl.Error:
            pass


def test_self_synchronising_decode_host_twin(vbl):
    """Scans without restart markers decoded by subsequences (csrc/jpeg.cu: blind starts, passes from the left
    neighbour's end state until nobody decodes again, prefix sum of block counts, write pass, DC scan): the same pixels
    as libjpeg-turbo for every subsequence size that lets the passes settle, and a loud failure when they are cut short."""
    import ctypes as C
    L = vbl.lib()
    PI = C.POINTER(C.c_int)
    L.vb200_debug_jpeg_decode_sync.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, PI, PI, PI, PI]

    def dec(d, shrink, sub, passes):
        w, h, b, used = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        vbl._check(L.vb200_debug_jpeg_decode_sync(d, len(d), shrink, sub, passes, None, 0, C.byref(w), C.byref(h), C.byref(b), C.byref(used)))
        out = np.zeros((h.value, w.value, b.value), np.uint8)
        vbl._check(L.vb200_debug_jpeg_decode_sync(d, len(d), shrink, sub, passes, out.ctypes.data_as(C.c_void_p), w.value * b.value,
                                                  C.byref(w), C.byref(h), C.byref(b), C.byref(used)))
        return out, used.value

    worst = 0
    for (h, w) in ((256, 320), (203, 301), (64, 64)):
        a = synth(h, w, seed=h)
        for sub in (2, 0):
            for q in (95, 75, 30):
                d = encode(a, q, sub, **({"optimize": True} if q == 75 else {}))
                for sb in (128, 256, 1024, 1 << 20):
                    got, used = dec(d, 2, sb, 4096)
                    assert np.array_equal(got, turbo_decode(d, 2)), (h, w, sub, q, sb)
                    worst = max(worst, used)
                    assert sb < (1 << 20) or used == 1
    assert worst >= 3          # the passes really were needed
    g = encode(synth(150, 203, seed=3, grey=True), 90)
    assert np.array_equal(dec(g, 1, 256, 4096)[0], turbo_decode(g, 1))
    # cut short: an error, never wrong pixels
    d = encode(synth(256, 320, seed=9), 100, 0)
    with pytest.raises(vbl.Error, match="converge|corrupt"):
        dec(d, 2, 128, 3)


def test_damaged_streams_never_fault(vbl):
    """bit flips, overwritten bytes and cuts anywhere in baseline / progressive / restart-marker streams: the decoder
    returns pixels or an error (jpeg2vips.c's fail_on decides which the caller wants), it never reads or writes outside"""
    rng = np.random.default_rng(123)
    a = synth(96, 120, seed=1)
    base = [encode(a, 70, sub, **kw) for kw in ({}, {"progressive": True}, {"restart_marker_rows": 1},
                                                 {"progressive": True, "restart_marker_rows": 1}, {"optimize": True}) for sub in (2, 1, 0)]
    decoded = rejected = 0
    for it in range(900):
        d = bytearray(base[it % len(base)])
        for _ in range(int(rng.integers(1, 6))):
            p = int(rng.integers(2, len(d)))
            mode = rng.integers(0, 3)
            if mode == 0:
                d[p] ^= 1 << int(rng.integers(0, 8))
            elif mode == 1:
                d[p] = int(rng.integers(0, 256))
            else:
                d = d[:p] + d[p + int(rng.integers(1, 4)):]
            if len(d) < 8:
                break
        try:
            out = vbl.jpeg_decode_host_twin(bytes(d), int(rng.choice([1, 2, 4, 8])))
            assert out.ndim == 3
            decoded += 1
        except vbl.Error:
            rejected += 1
    assert decoded > 100 and rejected > 100


def test_jpegshrink_rule(vbl):
    """thumbnail.c:489-517: shrink >= 16 -> 8, >= 8 -> 4, >= 4 -> 2, else 1, on the common shrink"""
    assert vbl.thumbnail_jpegshrink(4096, 4096, 512) == 4
    assert vbl.thumbnail_jpegshrink(4096, 4096, 256) == 8
    assert vbl.thumbnail_jpegshrink(4096, 4096, 1024) == 2
    assert vbl.thumbnail_jpegshrink(4096, 4096, 1025) == 1
    assert vbl.thumbnail_jpegshrink(6000, 4000, 300) == 8
    assert vbl.thumbnail_jpegshrink(6000, 4000, 300, 300, "force") == 4    # common = min(20, 13.3)
    assert vbl.thumbnail_jpegshrink(800, 600, 1000) == 1


def test_geometry_without_gpu(vbl):
    a = synth(203, 301)
    streams = [encode(a, 85, 2), encode(a[::-1].copy(), 60, 2)]
    assert vbl.jpeg_geometry(streams, 4) == (75, 50, 3)
    with pytest.raises(vbl.Error, match="geometry"):
        vbl.jpeg_geometry([streams[0], encode(a[:100], 85, 2)], 4)


# ------------------------------------------------------------------ GPU


@pytest.mark.gpu
def test_gpu_batch_decode_matches_libjpeg_turbo(vbl):
    import libvips_b200 as vb
    vb.init(0)
    for (h, w) in ((256, 320), (203, 301), (1024, 1024)):
        for sub in (2, 1, 0):
            for kw in ({}, {"restart_marker_rows": 1}, {"optimize": True}, {"progressive": True}, {"progressive": True, "restart_marker_rows": 2}):
                streams = [encode(synth(h, w, seed=i), (95, 75, 40)[i], sub, **kw) for i in range(3)]
                if kw.get("progressive") and sub != 1:
                    streams.append(encode(synth(h, w, seed=9), 85, sub))       # a baseline frame in the same batch (not always)
                for shrink in (8, 4, 2, 1):
                    got = vb.jpeg_decode_batch(streams, shrink)
                    want = np.stack([turbo_decode(s, shrink) for s in streams])
                    assert got.shape == want.shape and np.array_equal(got, want), (h, w, sub, kw, shrink)
    g = [encode(synth(150, 203, seed=i, grey=True), 90) for i in range(2)]
    assert np.array_equal(vb.jpeg_decode_batch(g, 2), np.stack([turbo_decode(s, 2) for s in g]))


@pytest.mark.gpu
def test_gpu_self_synchronising_decode(vbl):
    """streams without restart markers through the subsequence kernels: forced onto small images with small
    subsequences (VB200_JPEG_SYNC), and by default on frames big enough to take that path by themselves"""
    import os
    import libvips_b200 as vb
    vb.init(0)
    for sb in ("128", "512"):
        os.environ["VB200_JPEG_SYNC"] = sb
        try:
            for (h, w) in ((256, 320), (203, 301)):
                for sub in (2, 0):
                    streams = [encode(synth(h, w, seed=i), (95, 75, 40)[i], sub, **({"optimize": True} if i == 1 else {})) for i in range(3)]
                    streams.append(encode(synth(h, w, seed=7), 85, sub, restart_marker_rows=1))     # a DRI frame in the same batch
                    for shrink in (4, 2):
                        got = vb.jpeg_decode_batch(streams, shrink)
                        want = np.stack([turbo_decode(s, shrink) for s in streams])
                        assert np.array_equal(got, want), (sb, h, w, sub, shrink)
        finally:
            del os.environ["VB200_JPEG_SYNC"]
    big = [encode(synth(2048, 2048, seed=i), (92, 80)[i], 2) for i in range(2)]
    assert all(len(s) > 64 * 1024 for s in big)
    assert np.array_equal(vb.jpeg_decode_batch(big, 4), np.stack([turbo_decode(s, 4) for s in big]))
    os.environ["VB200_JPEG_SYNC"] = "0"         # one thread per frame: the same pixels
    try:
        assert np.array_equal(vb.jpeg_decode_batch(big[:1], 4), np.stack([turbo_decode(s, 4) for s in big[:1]]))
    finally:
        del os.environ["VB200_JPEG_SYNC"]


@pytest.mark.gpu
def test_gpu_corrupt_scan_is_an_error_or_decodes(vbl):
    import libvips_b200 as vb
    vb.init(0)
    d = bytearray(encode(synth(128, 128), 85, 2, restart_marker_rows=1))
    rng = np.random.default_rng(9)
    for _ in range(8):
        bad = bytearray(d)
        for p in rng.integers(700, len(d) - 2, 12):
            if bad[p] != 0xFF and bad[p - 1] != 0xFF:
                bad[p] ^= 1 << int(rng.integers(0, 8))
        try:
            out = vb.jpeg_decode_batch([bytes(bad)], 2)     # never a crash, never an out-of-bounds access
            assert out.shape == (1, 64, 64, 3)
        except vb.Error:
            pass


@pytest.mark.gpu
def test_gpu_thumbnail_from_jpeg_streams(vbl):
    """vips_thumbnail_buffer's chain: shrink-on-load picked by thumbnail.c:489-517, then the thumbnail of the decoded
    frame -- against the oracle thumbnail of libjpeg-turbo's own decode"""
    import libvips_b200 as vb
    from oracle import pyoracle
    vb.init(0)
    for (h, w, target) in ((2048, 2048, 256), (1536, 2048, 200), (1024, 1024, 200)):
        streams = [encode(synth(h, w, seed=i), 88, 2, **({"restart_marker_rows": 1} if i else {})) for i in range(3)]
        shrink = vb.thumbnail_jpegshrink(w, h, target)
        assert shrink in (2, 4, 8)
        dw, dh, bands = vb.jpeg_geometry(streams, shrink)
        assert (dw, dh, bands) == (w // shrink, h // shrink, 3)
        plan = vb.ThumbnailPlan(dw, dh, 3, target)
        got = plan.run_jpeg(streams, shrink)
        want = np.stack([pyoracle.thumbnail_image(turbo_decode(s, shrink), target) for s in streams])
        assert got.shape == want.shape and np.array_equal(got, want), (h, w, target)


@pytest.mark.gpu
def test_gpu_thumbnail_buffer(vbl):
    """vb200_thumbnail_buffer = vips_thumbnail_buffer for a JPEG stream: the load-time shrink of thumbnail.c:489-517 (1 for a
    modest reduction: full-size 4:2:0 through the fancy upsampler), then the thumbnail of what was loaded"""
    import libvips_b200 as vb
    from oracle import pyoracle
    vb.init(0)
    for (h, w, target, sub) in ((1024, 1536, 200, 2), (600, 800, 300, 2), (512, 512, 100, 1), (900, 700, 64, 0)):
        d = encode(synth(h, w, seed=target), 85, sub)
        shrink = vb.thumbnail_jpegshrink(w, h, target)
        got = vb.thumbnail_buffer(d, target)
        want = pyoracle.thumbnail_image(turbo_decode(d, shrink), target)
        assert got.shape == want.shape and np.array_equal(got, want), (h, w, target, sub, shrink)


@pytest.mark.gpu
def test_gpu_concurrent_callers(vbl):
    """libvips calls loaders from worker threads: the pump (pinned / device slots) is one per process, callers take turns"""
    import threading
    import libvips_b200 as vb
    vb.init(0)
    jobs = [([encode(synth(256 + 16 * t, 320, seed=10 * t + i), 80, (2, 1, 0)[t % 3]) for i in range(3)], (2, 1, 4, 2)[t]) for t in range(4)]
    want = [np.stack([turbo_decode(s, shrink) for s in streams]) for streams, shrink in jobs]
    got, errs = [None] * len(jobs), []

    def work(t):
        try:
            for _ in range(3):
                got[t] = vb.jpeg_decode_batch(jobs[t][0], jobs[t][1])
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs
    for t in range(len(jobs)):
        assert np.array_equal(got[t], want[t]), t

