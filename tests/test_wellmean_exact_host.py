"""The integer decomposition behind the exact well sums (csrc/wellmean_exact.cuh) compiled for the
host: every finite float32 must equal its (class, multiple) pair exactly, so class sums are plain
integer sums and the per-well means cannot depend on the order of the rows."""
import ctypes
import os
import shutil
import subprocess
from fractions import Fraction

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "image_processing_suite_b200", "csrc", "wellmean_exact.cuh")

SRC = r'''
#include "%s"
extern "C" int split(uint32_t bits, int* cls, long long* mult) { return wmx_split(bits, cls, mult) ? 1 : 0; }
extern "C" int unit_exp(int c) { return wmx_unit_exp(c); }
extern "C" void class_sums(const uint32_t* bits, long n, long long* acc) {
  for (long i = 0; i < n; ++i) { int c; long long m; if (wmx_split(bits[i], &c, &m)) acc[c] += m; }
}
extern "C" double total(const long long* acc) { return wmx_total(acc); }
''' % HDR


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    d = tmp_path_factory.mktemp("wmx")
    src, so = d / "wmx.cpp", d / "libwmx.so"
    src.write_text(SRC)
    subprocess.run([gxx, "-O2", "-shared", "-fPIC", "-x", "c++", str(src), "-o", str(so)], check=True)
    L = ctypes.CDLL(str(so))
    L.total.restype = ctypes.c_double
    return L


def exact(bits):
    return Fraction(float(np.array([bits], np.uint32).view(np.float32)[0]))


def test_every_kind_of_float32_splits_exactly(lib):
    rng = np.random.default_rng(0)
    specials = [0x00000000, 0x80000000, 0x00000001, 0x80000001, 0x007fffff, 0x00800000, 0x00ffffff, 0x03ffffff,
                0x04000000, 0x3f800000, 0xbf800000, 0x7f7fffff, 0xff7fffff, 0x7c000000, 0x7bffffff]
    bits = np.concatenate([np.array(specials, np.uint32), rng.integers(0, 2 ** 32, 20000, dtype=np.uint64).astype(np.uint32)])
    bits = bits[((bits >> 23) & 255) != 255]                    # finite only (the kernel flags inf, skips NaN)
    for b in bits.tolist():
        c, m = ctypes.c_int(), ctypes.c_longlong()
        nz = lib.split(ctypes.c_uint32(b), ctypes.byref(c), ctypes.byref(m))
        assert 0 <= c.value < 32 and abs(m.value) < 2 ** 31
        assert Fraction(m.value) * Fraction(2) ** lib.unit_exp(c.value) == exact(b)
        assert bool(nz) == (exact(b) != 0)
        assert c.value == max((b >> 23) & 255, 1) >> 3


def test_class_sums_are_exact_and_order_independent(lib):
    rng = np.random.default_rng(1)
    vals = (rng.normal(0.0, 1.0, 50000) * 10.0 ** rng.integers(-30, 30, 50000)).astype(np.float32)
    vals[:100] = np.float32(1e-42)                               # denormals
    bits = np.ascontiguousarray(vals.view(np.uint32))
    acc = (ctypes.c_longlong * 32)()
    lib.class_sums(bits.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(bits.size), acc)
    want = sum((Fraction(float(v)) for v in vals), Fraction(0))
    got = sum((Fraction(acc[c]) * Fraction(2) ** lib.unit_exp(c) for c in range(32)), Fraction(0))
    assert got == want
    perm = np.ascontiguousarray(bits[rng.permutation(bits.size)])
    acc2 = (ctypes.c_longlong * 32)()
    lib.class_sums(perm.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(perm.size), acc2)
    assert list(acc) == list(acc2)
    t = lib.total(acc)
    assert abs(Fraction(t) - want) <= abs(want) * Fraction(1, 2 ** 48)
    # a narrow column (what an object feature looks like): the float64 total is the rounded exact sum
    area = rng.integers(80, 900, 18000).astype(np.float32)
    acc3 = (ctypes.c_longlong * 32)()
    lib.class_sums(area.view(np.uint32).ctypes.data_as(ctypes.c_void_p), ctypes.c_long(area.size), acc3)
    assert lib.total(acc3) == float(area.astype(np.float64).sum())
