// Exact, order-independent sums of float32 values in integer arithmetic (well aggregation).
//
// A finite float32 is M * 2^(max(E,1) - 150) with a 24-bit integer M (E = biased exponent, the
// implicit bit included for E >= 1).  Values whose exponents share E >> 3 form a "class" of 8
// binades; inside class c every value is an integer multiple, below 2^31, of the class unit
// 2^(max(8c,1) - 150).  A class sum is therefore a plain 64-bit integer sum: associative, so
// independent of the order in which threads, chunks and ranks deliver the rows, and native as a
// fire-and-forget 64-bit integer atomic.  Only the final sum over the 32 classes (ascending, in
// float64, by one thread) rounds.  Compiles for the host too (tests/test_wellmean_exact_host.py
// checks it against exact rational arithmetic).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define WMX_HD __host__ __device__ __forceinline__
#else
#define WMX_HD static inline
#endif

#define WMX_CLASSES 32

// power-of-two exponent of the unit of class c
WMX_HD int wmx_unit_exp(int c) { return (c ? 8 * c : 1) - 150; }

// bits of a FINITE float32 -> (class, signed multiple of the class unit); false for +-0.
WMX_HD bool wmx_split(uint32_t bits, int* cls, long long* mult) {
  uint32_t E = (bits >> 23) & 255u;
  uint32_t M = bits & 0x7fffffu;
  if (E) M |= 0x800000u; else E = 1u;          // denormals share the unit of E = 1
  const uint32_t c = E >> 3;
  const uint32_t sh = E - (c ? 8u * c : 1u);   // 0 .. 7
  const long long v = (long long)((uint64_t)M << sh);
  *cls = (int)c;
  *mult = (bits >> 31) ? -v : v;
  return M != 0u;
}

// 2^e as a float64, e in [-1022, 1023]
WMX_HD double wmx_pow2(int e) {
  union { uint64_t u; double d; } x;
  x.u = (uint64_t)(1023 + e) << 52;
  return x.d;
}

// the float64 value of the 32 class sums of one (well, column): ascending classes, small terms first
WMX_HD double wmx_total(const long long* acc) {
  double s = 0.0;
  for (int c = 0; c < WMX_CLASSES; ++c)
    if (acc[c]) s += (double)acc[c] * wmx_pow2(wmx_unit_exp(c));
  return s;
}

// ---- what a thread carries for the (well, column) it is summing -----------------------------------
// The sums of the two exponent classes seen last (a column's values rarely straddle more), held as
// float64: at most 2^16 values of one class are integers below 2^47 in units of the class -- exact.
// They become integer multiples of the class unit when they leave through `sink(class, units)`.
struct WmxSlots {
  double a0 = 0.0, a1 = 0.0;
  int c0 = -1, c1 = -1, nan = 0, inf = 0;
};

WMX_HD int wmx_class(uint32_t bits) { return (int)((bits >> 26) & 31u); }

// a class sum held as float64 -> its (exact) integer multiple of the class unit
WMX_HD unsigned long long wmx_units(double a, int c) {
  return (unsigned long long)(long long)(a * wmx_pow2(-wmx_unit_exp(c)));
}

// Opening the slots ahead of the values: called on the first few values of a run, in order, it gives the
// two slots the first two classes that occur (zeros and the top class never claim one), so that the
// first values do not all pass through wmx_slow.
WMX_HD void wmx_seed(WmxSlots& t, uint32_t bits) {
  const int c = wmx_class(bits);
  const bool ok = (bits & 0x7fffffffu) != 0u && c != WMX_CLASSES - 1;
  if (ok && t.c0 < 0) t.c0 = c;
  else if (ok && c != t.c0 && t.c1 < 0) t.c1 = c;
}

// Hot path, branch-free (the lanes of a warp are different columns in different classes: an if-chain
// would run every arm for every warp): x = the value widened; it is added to the slot whose class it
// has, zero to the other.  Returns true when NO slot took it: wmx_slow must see the value.
WMX_HD bool wmx_fast(WmxSlots& t, uint32_t bits, double x) {
  const int c = wmx_class(bits);
  const bool h0 = c == t.c0, h1 = c == t.c1;
  t.a0 += h0 ? x : 0.0;
  t.a1 += h1 ? x : 0.0;
  return !(h0 || h1);
}

// Everything else: zero, NaN (skipped, counted in t.nan), +-inf (flag bits 1 / 2), the top class (never
// held in a slot: NaN and inf share its number), a class no slot holds (the older slot leaves).  Safe to
// call for any value, also one a slot has meanwhile been opened for.
template <class Sink>
WMX_HD void wmx_slow(WmxSlots& t, uint32_t bits, double x, Sink&& sink) {
  const int c = wmx_class(bits);
  const uint32_t absb = bits & 0x7fffffffu;
  if (absb == 0u) return;
  if (absb >= 0x7f800000u) {
    if (absb > 0x7f800000u) ++t.nan;
    else t.inf |= (bits >> 31) ? 2 : 1;
    return;
  }
  if (c == t.c0) { t.a0 += x; return; }
  if (c == t.c1) { t.a1 += x; return; }
  if (c == WMX_CLASSES - 1) { sink(c, wmx_units(x, c)); return; }
  if (t.c1 >= 0 && t.a1 != 0.0) sink(t.c1, wmx_units(t.a1, t.c1));
  t.c1 = t.c0; t.a1 = t.a0;
  t.c0 = c; t.a0 = x;
}

template <class Sink>
WMX_HD void wmx_leave(const WmxSlots& t, Sink&& sink) {
  if (t.c0 >= 0 && t.a0 != 0.0) sink(t.c0, wmx_units(t.a0, t.c0));
  if (t.c1 >= 0 && t.a1 != 0.0) sink(t.c1, wmx_units(t.a1, t.c1));
}
