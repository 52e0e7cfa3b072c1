"""CPU: the C-ABI library loads and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"^[A-Za-z_][\w \*]*?\b(vb200_\w+|vips_\w+_hwy)\s*\(", src, flags=re.M):
            names.append(m.group(1))
    return sorted(set(names))


def test_header_declares_something():
    names = declared_functions()
    assert "vb200_thumbnail_batch_device" in names and "vips_reducev_uchar_hwy" in names
    assert len(names) > 25


def test_library_exports_every_declared_symbol():
    import libvips_b200 as vb
    assert os.path.exists(vb.library_path()), "libvb200.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(vb.library_path())
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_error_buffer_without_gpu():
    """No compute here: only the error convention (0 / -1 + text buffer)."""
    import libvips_b200 as vb
    L = vb.lib()
    L.vb200_error_clear()
    assert L.vb200_error_buffer() == b""
    assert L.vb200_format_sizeof(0) == 1 and L.vb200_format_sizeof(6) == 4


def test_isenabled_switch(monkeypatch):
    """vb200_isenabled mirrors vips_vector_isenabled / VIPS_NOVECTOR (iofuncs/vector.cpp:98-110): VB200_DISABLE switches the
    device path off; otherwise it is on exactly when a CUDA device is present (none in the CPU container)"""
    import libvips_b200 as vb
    L = vb.lib()
    monkeypatch.setenv("VB200_DISABLE", "1")
    assert L.vb200_isenabled() == 0
    monkeypatch.setenv("VB200_DISABLE", "0")
    assert L.vb200_isenabled() in (0, 1)
    monkeypatch.delenv("VB200_DISABLE")
    assert L.vb200_isenabled() in (0, 1)
