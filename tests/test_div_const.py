"""The colour kernels divide by constants with a 3-instruction FMA sequence instead of a
full double division (colour.cu div_const).  This is the proof obligation: on the CPU, the
same sequence (glibc fma is exact) equals x / d bit for bit for every float mantissa and
for random doubles, for every divisor used (LabQ2sRGB.c / XYZ2Lab.c / Lab2XYZ.c constants)."""
import os
import subprocess


def test_constant_division_is_correctly_rounded(tmp_path):
    src = os.path.join(os.path.dirname(__file__), "div_const_check.c")
    exe = str(tmp_path / "div_const_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert r.stdout.count("mismatches 0") == 10, r.stdout
