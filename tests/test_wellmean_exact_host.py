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
#include <string.h>
#include "%s"
extern "C" int split(uint32_t bits, int* cls, long long* mult) { return wmx_split(bits, cls, mult) ? 1 : 0; }
extern "C" int unit_exp(int c) { return wmx_unit_exp(c); }
extern "C" void class_sums(const uint32_t* bits, long n, long long* acc) {
  for (long i = 0; i < n; ++i) { int c; long long m; if (wmx_split(bits[i], &c, &m)) acc[c] += m; }
}
extern "C" double total(const long long* acc) { return wmx_total(acc); }
// one thread of the kernel: batches of 8 through the branch-free adds, the values no slot took seen again
// by wmx_slow after the batch (wm_batch in wellmean.cu), the tail value by value, slots leave at the end
extern "C" void thread_sums(const float* v, long n, long long* acc, int* nan_inf) {
  WmxSlots t;
  auto sink = [&](int c, unsigned long long units) { acc[c] += (long long)units; };
  long i = 0;
  for (; i + 8 <= n; i += 8) {
    if (i == 0) for (int j = 0; j < 8; ++j) { uint32_t b; memcpy(&b, &v[j], 4); wmx_seed(t, b); }
    unsigned odd = 0u;
    for (int j = 0; j < 8; ++j) { uint32_t b; memcpy(&b, &v[i + j], 4); odd |= wmx_fast(t, b, (double)v[i + j]) ? (1u << j) : 0u; }
    if (odd) for (int j = 0; j < 8; ++j) if (odd & (1u << j)) { uint32_t b; memcpy(&b, &v[i + j], 4); wmx_slow(t, b, (double)v[i + j], sink); }
  }
  for (; i < n; ++i) { uint32_t b; memcpy(&b, &v[i], 4); if (wmx_fast(t, b, (double)v[i])) wmx_slow(t, b, (double)v[i], sink); }
  wmx_leave(t, sink);
  nan_inf[0] = t.nan; nan_inf[1] = t.inf;
}
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


@pytest.mark.parametrize("kind", ["one class", "two classes", "many classes", "specials"])
def test_thread_slots_deliver_the_exact_class_sums(lib, kind):
    """The per-thread part of the kernel (two float64 class slots, deferred rare path) hands exactly the
    class sums of its values to the integer accumulators, whatever the mix of classes; NaNs are counted,
    infinities flagged, neither reaches a sum."""
    rng = np.random.default_rng({"one class": 3, "two classes": 4, "many classes": 5, "specials": 6}[kind])
    n = 64
    if kind == "one class":
        v = rng.uniform(600.0, 900.0, n)
    elif kind == "two classes":
        v = rng.uniform(80.0, 900.0, n) * rng.choice([-1.0, 1.0], n)        # straddles 512 = a class boundary
    elif kind == "many classes":
        v = rng.normal(0.0, 1.0, n) * 10.0 ** rng.integers(-30, 30, n)
    else:
        v = rng.normal(0.0, 1.0, n) * 10.0 ** rng.integers(-3, 3, n)
        v[[1, 9, 10, 33]] = np.nan
        v[[2, 40]] = np.inf
        v[5] = -np.inf
        v[[0, 7, 8, 63]] = 0.0
        v[11] = -0.0
        v[12] = 3.0e38                                                      # the top class: never held in a slot
        v[13] = 1.0e-42                                                     # a denormal
    v = v.astype(np.float32)
    for m in (n, n - 3, 5):                                                 # full batches, a tail, tail only
        part = np.ascontiguousarray(v[:m])
        acc = (ctypes.c_longlong * 32)()
        flags = (ctypes.c_int * 2)()
        lib.thread_sums(part.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(m), acc, flags)
        finite = np.ascontiguousarray(part[np.isfinite(part)])
        want = (ctypes.c_longlong * 32)()
        lib.class_sums(finite.view(np.uint32).ctypes.data_as(ctypes.c_void_p), ctypes.c_long(finite.size), want)
        assert list(acc) == list(want)
        assert flags[0] == int(np.isnan(part).sum())
        assert flags[1] == (1 if np.any(part == np.inf) else 0) | (2 if np.any(part == -np.inf) else 0)
