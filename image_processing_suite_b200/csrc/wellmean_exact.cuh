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
