// K1 + K3 in one pass, second generation (the benchmarked kernel): the same arithmetic as
// field_fused.cu -- z-max projection -> illumination divide -> b x b sum bin, and the label-keyed
// per-object reduction (object_accum.cuh: lane -> warp tree -> CTA records -> global atomics) --
// on an instruction diet.  The round-1 kernel was issue-bound (572 M warp instructions per 16
// fields, 0.79 of the HBM peak); measured on B200 (tools/ubench/f32x2.cu) the ALU pipe (LOP3,
// IADD3, PRMT, FMNMX, ISETP, SEL) takes one warp instruction every other cycle per scheduler and
// the FMA pipe one per cycle, so what counts is issue slots first and ALU-pipe work second:
//   * label masks travel and are analysed as packed uint16 pairs (Cellpose's own dtype): one
//     128-bit load per row; slot analysis is branch-free on the pairs (packed max, xor, packed
//     min against 0x00010001) instead of a per-pixel if-chain (15 % of the old kernel's
//     instructions were control flow);
//   * all per-pixel float work is issued as packed pairs (fma/add/mul.rn.f32x2 -> FFMA2 / FADD2:
//     two results per issue slot); uint16 -> float is a byte permute into 2^23 + v followed by a
//     packed subtract, no I2F;
//   * the illumination function may be passed as its reciprocal (ips_illum_reciprocal, once per
//     plate): the divide is then one packed multiply, no MUFU;
//   * moments: masked min / max first, the RUN's minimum (segmented tree over the lanes that
//     share the label) is the pivot of every lane of the run, so sum(q - p) and sum((q - p)^2)
//     go through the shuffle tree as float32 and become float64 once per run, not once per lane;
//   * geometry rides the tree in four packed words (area | sum y, sum x, packed min, packed max)
//     relative to the warp's origin instead of seven full-width values;
//   * addresses are one 64-bit base per channel plus 32-bit offsets.
// Results: maxproj and integer features bit-exact, float features within the 1e-5 contract
// (tests/test_gpu_field_fused.py compares with the oracle at full size).
//
// Replaces np.maximum.reduce (MaxProjection.py:45), img.astype(float)/illum
// (Illumination_QC_mult.py:145-150, Cellpose_GPU_s3fs.py:72), north_star's sum re-binning and
// the CellProfiler MeasureObject* subprocess (Feature_extraction_opt.py:166-167).
#include <stdlib.h>

#include "object_accum.cuh"

namespace ips {

typedef unsigned long long u64;

// ---- packed float32 pairs --------------------------------------------------------------------
__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ u64 pk2u(unsigned lo, unsigned hi) {
  u64 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float rcp_approx(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r;
}

constexpr int F2_WPR = OA_PX / 2;   // label / pixel words (uint16 pairs) per row of a lane

// ---- the per-lane state that lives across the channel loop --------------------------------------
template <int ROWS>
struct F2Lane {
  unsigned L1, L2;          // slot labels (0: none)
  unsigned m1, m2, m3;      // pixel masks, bit r * 8 + i
  int rec1, rec2;           // record index in shared memory, -1: flush directly
  bool lead1, has2;
  int lane, seg_last, head_lane;
  bool t1, t2, t4;          // tree step d reaches a lane of the same run
  unsigned run_area;        // on run heads: pixels of the run's slot 1
  u64 nnw[ROWS][F2_WPR];    // -0.0f inside slot 1, -1.0f outside (packed pairs)
};

__device__ __forceinline__ unsigned half_max(unsigned w) { return max(w & 0xffffu, w >> 16); }

// e[k]: bit 0 = flag of pixel 2k, bit 16 = flag of pixel 2k + 1  ->  8 flags in pixel order
__device__ __forceinline__ unsigned row_mask(const unsigned (&e)[F2_WPR]) {
  unsigned t = e[1] * 4u + e[0];
  t = e[2] * 16u + t;
  t = e[3] * 64u + t;
  return (t | (t >> 15)) & 0xffu;
}

// Branch-free label analysis of a ROWS x 8 window held as uint16 pairs.  Slot 1 = the LARGEST
// label of the window, slot 2 = the largest of the rest, m3 = anything else (rare: flushed per
// pixel).  Which label gets which slot does not matter to the results.
template <int ROWS>
__device__ __forceinline__ void f2_analyze(unsigned (&w)[ROWS][F2_WPR], unsigned Nmax, F2Lane<ROWS>& L,
                                           bool& overflow) {
  unsigned mx = 0u;
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int k = 0; k < F2_WPR; ++k) mx = __vmaxu2(mx, w[r][k]);
  unsigned L1 = half_max(mx);
  if (L1 > Nmax) {   // rare: labels beyond the caller's bound are reported and skipped
    overflow = true;
    mx = 0u;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int k = 0; k < F2_WPR; ++k) {
        unsigned v = w[r][k];
        if ((v & 0xffffu) > Nmax) v &= 0xffff0000u;
        if ((v >> 16) > Nmax) v &= 0x0000ffffu;
        w[r][k] = v;
        mx = __vmaxu2(mx, v);
      }
    L1 = half_max(mx);
  }
  L.L1 = L1;
  const unsigned rep1 = L1 ? L1 * 0x10001u : 0xffffffffu;   // no pixel matches when the window is background
  unsigned rest[ROWS][F2_WPR];
  unsigned m1 = 0u, mr = 0u;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    unsigned e[F2_WPR];
#pragma unroll
    for (int k = 0; k < F2_WPR; ++k) {
      const unsigned nz = __vminu2(w[r][k] ^ rep1, 0x00010001u);   // 1 per half that is NOT slot 1
      e[k] = nz ^ 0x00010001u;
      rest[r][k] = w[r][k] & (nz * 0xffffu);
      mr = __vmaxu2(mr, rest[r][k]);
      L.nnw[r][k] = pk2u((nz & 1u) * 0x3f800000u + 0x80000000u, (nz >> 16) * 0x3f800000u + 0x80000000u);
    }
    m1 |= row_mask(e) << (r * OA_PX);
  }
  L.m1 = m1;
  const unsigned L2 = half_max(mr);
  L.L2 = L2;
  L.has2 = L2 != 0u;
  unsigned m2 = 0u, m3 = 0u;
  if (__any_sync(OA_FULL, L.has2)) {
    const unsigned rep2 = L2 ? L2 * 0x10001u : 0xffffffffu;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      unsigned e2[F2_WPR], e3[F2_WPR];
#pragma unroll
      for (int k = 0; k < F2_WPR; ++k) {
        const unsigned nz2 = __vminu2(rest[r][k] ^ rep2, 0x00010001u);
        e2[k] = nz2 ^ 0x00010001u;
        e3[k] = __vminu2(rest[r][k] & (nz2 * 0xffffu), 0x00010001u);
      }
      m2 |= row_mask(e2) << (r * OA_PX);
      m3 |= row_mask(e3) << (r * OA_PX);
    }
  }
  L.m2 = m2;
  L.m3 = m3;
}

// Geometry of slot 1 relative to the warp's origin (row y0, column xw0 = first column of lane 0),
// packed for the tree: A = area | sum(y - y0) << 16, B = sum(x - xw0), MN = xmin | ymin << 16,
// MX = xmax1 | ymax1 << 16 (empty: 0xffff / 0).
template <int ROWS>
__device__ __forceinline__ void f2_geom_packed(unsigned m, int lane, unsigned& A, unsigned& B, unsigned& MN,
                                               unsigned& MX) {
  const unsigned area = __popc(m);
  unsigned sy = 0u;
#pragma unroll
  for (int r = 1; r < ROWS; ++r) sy += (unsigned)r * __popc((m >> (r * OA_PX)) & 0xffu);
  const unsigned colsum = __popc(m & 0xaaaaaaaau) + 2u * __popc(m & 0xccccccccu) + 4u * __popc(m & 0xf0f0f0f0u);
  unsigned cols = m;
  if (ROWS > 2) cols |= cols >> 16;
  if (ROWS > 1) cols |= cols >> 8;
  cols &= 0xffu;
  const unsigned xl = (unsigned)lane * OA_PX;
  A = area | (sy << 16);
  B = xl * area + colsum;
  if (m) {
    MN = (xl + __ffs(cols) - 1) | ((unsigned)((__ffs(m) - 1) >> 3) << 16);
    MX = (xl + 32 - __clz(cols)) | ((unsigned)(((31 - __clz(m)) >> 3) + 1) << 16);
  } else {
    MN = 0xffffffffu;
    MX = 0u;
  }
}

#define F2_TREE(v, op)                                              \
  do {                                                              \
    auto o1__ = __shfl_down_sync(OA_FULL, v, 1);                    \
    if (L.t1) v = op(v, o1__);                                      \
    auto o2__ = __shfl_down_sync(OA_FULL, v, 2);                    \
    if (L.t2) v = op(v, o2__);                                      \
    auto o4__ = __shfl_down_sync(OA_FULL, v, 4);                    \
    if (L.t4) v = op(v, o4__);                                      \
  } while (0)
#define F2_ADD(a, b) ((a) + (b))

// Call once per lane after the labels are loaded (all 32 lanes of the warp must call).
template <int ROWS, typename SH>
__device__ __forceinline__ void f2_begin(F2Lane<ROWS>& L, unsigned (&w)[ROWS][F2_WPR], unsigned Nmax, int y0, int xw0,
                                         SH& sh, unsigned long long* __restrict__ rec_f, int C,
                                         bool& overflow) {
  L.lane = threadIdx.x & 31;
  f2_analyze<ROWS>(w, Nmax, L, overflow);
  // runs of equal slot-1 labels along the warp, cut every OA_SEG lanes
  const unsigned prev = __shfl_up_sync(OA_FULL, L.L1, 1);
  const bool head = ((L.lane & (OA_SEG - 1)) == 0) || (prev != L.L1);
  const unsigned heads = __ballot_sync(OA_FULL, head);
  const unsigned above = heads & ~((2u << L.lane) - 1u);
  L.seg_last = above ? (__ffs(above) - 2) : 31;
  L.head_lane = 31 - __clz(heads & ((2u << L.lane) - 1u));
  L.t1 = L.lane + 1 <= L.seg_last;
  L.t2 = L.lane + 2 <= L.seg_last;
  L.t4 = L.lane + 4 <= L.seg_last;
  L.lead1 = head && L.L1 != 0u;
  const unsigned need1 = __ballot_sync(OA_FULL, L.lead1);
  const unsigned need2 = __ballot_sync(OA_FULL, L.has2);
  const int n1 = __popc(need1), n2 = __popc(need2);
  int base = 0;
  if (L.lane == 0 && n1 + n2 > 0) base = atomicAdd(&sh.count, n1 + n2);
  base = __shfl_sync(OA_FULL, base, 0);
  const unsigned lt = (1u << L.lane) - 1u;
  L.rec1 = base + __popc(need1 & lt);
  L.rec2 = base + n1 + __popc(need2 & lt);
  if (!L.lead1 || L.rec1 >= SH::CAP) L.rec1 = -1;
  if (!L.has2 || L.rec2 >= SH::CAP) L.rec2 = -1;

  // geometry: slot 1 through the tree in packed form, slot 2 and the remainder directly
  unsigned A, B, MN, MX;
  f2_geom_packed<ROWS>(L.m1, L.lane, A, B, MN, MX);
  F2_TREE(A, F2_ADD);
  F2_TREE(B, F2_ADD);
  F2_TREE(MN, __vminu2);
  F2_TREE(MX, __vmaxu2);
  L.run_area = A & 0xffffu;
  if (L.lead1) {
    OaGeom g;
    g.area = L.run_area;
    g.sy = (unsigned)y0 * g.area + (A >> 16);
    g.sx = (unsigned)xw0 * g.area + B;
    g.xmin = xw0 + (int)(MN & 0xffffu);
    g.ymin = y0 + (int)(MN >> 16);
    g.xmax1 = xw0 + (int)(MX & 0xffffu);
    g.ymax1 = y0 + (int)(MX >> 16);
    if (L.rec1 >= 0) { sh.label[L.rec1] = (int)L.L1; sh.geom[L.rec1] = g; }
    else oa_flush_geom(rec_f, C, (int)L.L1, g);
  }
  const int x0 = xw0 + L.lane * OA_PX;
  if (L.has2) {
    const OaGeom g2 = oa_geom<ROWS>(L.m2, y0, x0);
    if (L.rec2 >= 0) { sh.label[L.rec2] = (int)L.L2; sh.geom[L.rec2] = g2; }
    else oa_flush_geom(rec_f, C, (int)L.L2, g2);
  }
  if (L.m3) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int i = 0; i < OA_PX; ++i)
        if ((L.m3 >> (r * OA_PX + i)) & 1u)
          oa_flush_geom(rec_f, C, (int)((w[r][i >> 1] >> (16 * (i & 1))) & 0xffffu),
                        oa_geom<ROWS>(1u << (r * OA_PX + i), y0, x0));
  }
}

// Where the lane's labels live, for the rare third-label path (everything but `elem` is uniform).
struct F2Labels {
  const void* base;
  size_t elem;          // index of the window's first label
  int label_bytes, W;
};
// label of pixel j of the lane's window, re-read on the rare third-label path
__device__ __noinline__ unsigned f2_reload_label(const void* base, size_t elem, int label_bytes, int W, int j) {
  const size_t e = elem + (size_t)(j / OA_PX) * W + (j % OA_PX);
  return label_bytes == 2 ? (unsigned)reinterpret_cast<const uint16_t*>(base)[e]
                          : (unsigned)reinterpret_cast<const int32_t*>(base)[e];
}

// ---- one channel, float mode ---------------------------------------------------------------------
// q[r][k]: the corrected pixels of the lane's window as packed pairs.
template <int ROWS, typename SH>
__device__ __forceinline__ void f2_channel_float(const F2Lane<ROWS>& L, int c, const u64 (&q)[ROWS][F2_WPR],
                                                 SH& sh, unsigned long long* __restrict__ rec_f, int C,
                                                 const F2Labels& lab) {
  const float BIG = 3.0e38f;
  {
    // pass A: masked min / max.  nnw = -0 inside / -1 outside: q + nnw * (-BIG) and q + q * nnw.
    const u64 NBIG2 = pk2(-BIG, -BIG);
    // two independent min / max chains (the warp has few neighbours to hide a serial one behind)
    float lo = BIG, hi = 0.f, lo_b = BIG, hi_b = 0.f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int k = 0; k < F2_WPR; ++k) {
        float a0, a1, b0, b1;
        upk2(fma2(L.nnw[r][k], NBIG2, q[r][k]), a0, a1);
        upk2(fma2(q[r][k], L.nnw[r][k], q[r][k]), b0, b1);
        if ((r * F2_WPR + k) & 1) {
          lo_b = fmin3(lo_b, a0, a1);
          hi_b = fmax3(hi_b, b0, b1);
        } else {
          lo = fmin3(lo, a0, a1);
          hi = fmax3(hi, b0, b1);
        }
      }
    lo = fminf(lo, lo_b);
    hi = fmaxf(hi, hi_b);
    F2_TREE(lo, fminf);
    F2_TREE(hi, fmaxf);
    // pass B around the run's minimum
    const float p = __shfl_sync(OA_FULL, lo, L.head_lane);
    const u64 NP2 = pk2(-p, -p);
    u64 s1 = 0ull, s2 = 0ull, t1 = 0ull, t2 = 0ull;   // packed +0.0f pairs, two chains each
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int k = 0; k < F2_WPR; ++k) {
        const u64 d = add2(q[r][k], NP2);
        const u64 dm = fma2(d, L.nnw[r][k], d);
        if ((r * F2_WPR + k) & 1) {
          t1 = add2(t1, dm);
          t2 = fma2(dm, d, t2);
        } else {
          s1 = add2(s1, dm);
          s2 = fma2(dm, d, s2);
        }
      }
    s1 = add2(s1, t1);
    s2 = add2(s2, t2);
    float a1, a1b, a2, a2b;
    upk2(s1, a1, a1b);
    upk2(s2, a2, a2b);
    a1 += a1b;
    a2 += a2b;
    F2_TREE(a1, F2_ADD);
    F2_TREE(a2, F2_ADD);
    if (L.lead1) {
      const double n = (double)L.run_area, P = (double)lo, A1 = (double)a1;
      const u64 d1 = (u64)__double_as_longlong(fma(n, P, A1));
      const u64 d2 = (u64)__double_as_longlong(fma(n * P, P, fma(2.0 * P, A1, (double)a2)));
      const unsigned mn = __float_as_uint(lo), mxv = __float_as_uint(hi);
      if (L.rec1 >= 0) { sh.s1[L.rec1][c] = d1; sh.s2[L.rec1][c] = d2; sh.mn[L.rec1][c] = mn; sh.mx[L.rec1][c] = mxv; }
      else oa_flush_chan<true>(rec_f, C, (int)L.L1, c, d1, d2, mn, mxv);
    }
  }
  if (__any_sync(OA_FULL, L.has2 || L.m3 != 0u)) {
    float fv[ROWS][OA_PX];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int k = 0; k < F2_WPR; ++k) upk2(q[r][k], fv[r][2 * k], fv[r][2 * k + 1]);
    if (L.has2) {
      float lo = __int_as_float(0x7f800000), hi = 0.f;
#pragma unroll
      for (int j = 0; j < ROWS * OA_PX; ++j) {
        const bool in = (L.m2 >> j) & 1u;
        const float v = fv[j / OA_PX][j % OA_PX];
        lo = fminf(lo, in ? v : lo);
        hi = fmaxf(hi, in ? v : hi);
      }
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int j = 0; j < ROWS * OA_PX; ++j) {
        const bool in = (L.m2 >> j) & 1u;
        const float d = in ? fv[j / OA_PX][j % OA_PX] - lo : 0.f;
        a1 += d;
        a2 = fmaf(d, d, a2);
      }
      unsigned long long d1, d2;
      oa_to_double(L.m2, lo, a1, a2, d1, d2);
      const unsigned mn = __float_as_uint(lo), mxv = __float_as_uint(hi);
      if (L.rec2 >= 0) { sh.s1[L.rec2][c] = d1; sh.s2[L.rec2][c] = d2; sh.mn[L.rec2][c] = mn; sh.mx[L.rec2][c] = mxv; }
      else oa_flush_chan<true>(rec_f, C, (int)L.L2, c, d1, d2, mn, mxv);
    }
    if (L.m3) {
#pragma unroll
      for (int j = 0; j < ROWS * OA_PX; ++j)
        if ((L.m3 >> j) & 1u) {
          const float v = fv[j / OA_PX][j % OA_PX];
          const double vd = (double)v;
          oa_flush_chan<true>(rec_f, C, (int)f2_reload_label(lab.base, lab.elem, lab.label_bytes, lab.W, j), c,
                              (u64)__double_as_longlong(vd), (u64)__double_as_longlong(vd * vd),
                              __float_as_uint(v), __float_as_uint(v));
        }
    }
  }
}

// ---- one channel, integer mode (no illumination function): exact sums ---------------------------
template <int ROWS>
__device__ __forceinline__ void f2_slot_int(unsigned m, const unsigned (&iv)[ROWS][OA_PX], u64& s1, u64& s2,
                                            unsigned& mn, unsigned& mx) {
  unsigned a1 = 0u, lo = 0xffffffffu, hi = 0u;
  u64 a2 = 0ull;
#pragma unroll
  for (int j = 0; j < ROWS * OA_PX; ++j) {
    const bool in = (m >> j) & 1u;
    const unsigned x = iv[j / OA_PX][j % OA_PX];
    const unsigned xm = in ? x : 0u;
    a1 += xm;
    a2 += (u64)(xm * xm);
    lo = min(lo, in ? x : lo);
    hi = max(hi, in ? x : hi);
  }
  s1 = a1; s2 = a2; mn = lo; mx = hi;
}

template <int ROWS, typename SH>
__device__ __forceinline__ void f2_channel_int(const F2Lane<ROWS>& L, int c, const unsigned (&iv)[ROWS][OA_PX],
                                               SH& sh, unsigned long long* __restrict__ rec_f, int C,
                                               const F2Labels& lab) {
  {
    u64 s1, s2;
    unsigned mn, mx;
    f2_slot_int<ROWS>(L.m1, iv, s1, s2, mn, mx);
    F2_TREE(s1, F2_ADD);
    F2_TREE(s2, F2_ADD);
    F2_TREE(mn, min);
    F2_TREE(mx, max);
    if (L.lead1) {
      if (L.rec1 >= 0) { sh.s1[L.rec1][c] = s1; sh.s2[L.rec1][c] = s2; sh.mn[L.rec1][c] = mn; sh.mx[L.rec1][c] = mx; }
      else oa_flush_chan<false>(rec_f, C, (int)L.L1, c, s1, s2, mn, mx);
    }
  }
  if (__any_sync(OA_FULL, L.has2 || L.m3 != 0u)) {
    if (L.has2) {
      u64 s1, s2;
      unsigned mn, mx;
      f2_slot_int<ROWS>(L.m2, iv, s1, s2, mn, mx);
      if (L.rec2 >= 0) { sh.s1[L.rec2][c] = s1; sh.s2[L.rec2][c] = s2; sh.mn[L.rec2][c] = mn; sh.mx[L.rec2][c] = mx; }
      else oa_flush_chan<false>(rec_f, C, (int)L.L2, c, s1, s2, mn, mx);
    }
    if (L.m3) {
#pragma unroll
      for (int j = 0; j < ROWS * OA_PX; ++j)
        if ((L.m3 >> j) & 1u) {
          const unsigned x = iv[j / OA_PX][j % OA_PX];
          oa_flush_chan<false>(rec_f, C, (int)f2_reload_label(lab.base, lab.elem, lab.label_bytes, lab.W, j), c,
                               (u64)x, (u64)(x * x), x, x);
        }
    }
  }
}

// ---- pieces shared by the two kernels ---------------------------------------------------------------
// Label words of the lane's window (uint16 pairs); int32 masks are narrowed here: labels <= 0
// are background, labels > Nmax (<= 65535) are reported and skipped.
template <int BIN>
__device__ __forceinline__ void f2_load_labels(unsigned (&lw)[BIN][F2_WPR], const char* lab_lane, int label_bytes,
                                               unsigned row_b, int Nmax, bool col_ok, uint64_t pol, bool& overflow) {
#pragma unroll
  for (int r = 0; r < BIN; ++r) {
    if (label_bytes == 2) {
      const uint4 l = ldg128_stream(lab_lane + (size_t)r * row_b, pol);
      lw[r][0] = l.x; lw[r][1] = l.y; lw[r][2] = l.z; lw[r][3] = l.w;
    } else {
      const uint4 l0 = ldg128_stream(lab_lane + (size_t)r * row_b * 2, pol);
      const uint4 l1 = ldg128_stream(lab_lane + (size_t)r * row_b * 2 + 16, pol);
      const unsigned v[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
      unsigned h[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool ok = v[i] <= (unsigned)Nmax;
        overflow |= col_ok && !ok && (int)v[i] > 0;
        h[i] = ok ? v[i] : 0u;
      }
#pragma unroll
      for (int k = 0; k < F2_WPR; ++k) lw[r][k] = h[2 * k] | (h[2 * k + 1] << 16);
    }
  }
#pragma unroll
  for (int r = 0; r < BIN; ++r)
#pragma unroll
    for (int k = 0; k < F2_WPR; ++k)
      if (!col_ok) lw[r][k] = 0u;      // lanes beyond the image width read a valid column group but own nothing
}

// Everything that happens to one channel of the lane's window once its pixels (m: packed uint16
// max projection rows) and illumination values (il: function or reciprocal) are in registers:
// divide, bin, store the binned row, fold the channel into the object records.
template <int BIN, bool HAS_ILLUM, typename SH>
__device__ __forceinline__ void f2_consume(const uint4 (&m)[BIN], const uint4 (&il)[BIN][2], int illum_is_rcp,
                                           bool store_ok, char* bp, uint64_t pol_stream, const F2Lane<BIN>& L,
                                           bool warp_fg, int c, SH& sh, unsigned long long* __restrict__ rec_f, int C,
                                           const F2Labels& lab) {
  constexpr int NB = OA_PX / BIN;
  if (HAS_ILLUM) {
    const u64 NMAGIC2 = pk2(-8388608.f, -8388608.f);
    u64 q[BIN][F2_WPR];
#pragma unroll
    for (int r = 0; r < BIN; ++r) {
      const unsigned mw[4] = {m[r].x, m[r].y, m[r].z, m[r].w};
      unsigned rc[8] = {il[r][0].x, il[r][0].y, il[r][0].z, il[r][0].w, il[r][1].x, il[r][1].y, il[r][1].z, il[r][1].w};
      if (!illum_is_rcp) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rc[i] = __float_as_uint(rcp_approx(__uint_as_float(rc[i])));
      }
#pragma unroll
      for (int k = 0; k < F2_WPR; ++k) {
        // 2^23 + v as float bits, then - 2^23: exact uint16 -> float without I2F
        const u64 fl = pk2u(__byte_perm(mw[k], 0x4b000000u, 0x7610), __byte_perm(mw[k], 0x4b000000u, 0x7632));
        q[r][k] = mul2(add2(fl, NMAGIC2), pk2u(rc[2 * k], rc[2 * k + 1]));
      }
    }
    if (store_ok) {
      float bs[NB];
      if (BIN == 1) {
#pragma unroll
        for (int k = 0; k < F2_WPR; ++k) upk2(q[0][k], bs[(2 * k) % NB], bs[(2 * k + 1) % NB]);
      } else {
        u64 vs[F2_WPR];
#pragma unroll
        for (int k = 0; k < F2_WPR; ++k) {
          vs[k] = q[0][k];
#pragma unroll
          for (int r = 1; r < BIN; ++r) vs[k] = add2(vs[k], q[r][k]);
        }
        float h[F2_WPR];
#pragma unroll
        for (int k = 0; k < F2_WPR; ++k) {
          float a, b;
          upk2(vs[k], a, b);
          h[k] = a + b;
        }
        if (BIN == 2) {
#pragma unroll
          for (int k = 0; k < F2_WPR; ++k) bs[k % NB] = h[k];
        } else {
          bs[0] = h[0] + h[1];
          bs[1 % NB] = h[2] + h[3];
        }
      }
      if (NB == 8) {
        stg128_stream(bp, make_uint4(__float_as_uint(bs[0]), __float_as_uint(bs[1 % NB]), __float_as_uint(bs[2 % NB]),
                                     __float_as_uint(bs[3 % NB])), pol_stream);
        stg128_stream(bp + 16, make_uint4(__float_as_uint(bs[4 % NB]), __float_as_uint(bs[5 % NB]),
                                          __float_as_uint(bs[6 % NB]), __float_as_uint(bs[7 % NB])), pol_stream);
      } else if (NB == 4) {
        stg128_stream(bp, make_uint4(__float_as_uint(bs[0]), __float_as_uint(bs[1 % NB]), __float_as_uint(bs[2 % NB]),
                                     __float_as_uint(bs[3 % NB])), pol_stream);
      } else {
        stg64_stream(bp, make_uint2(__float_as_uint(bs[0]), __float_as_uint(bs[1 % NB])), pol_stream);
      }
    }
    if (warp_fg) f2_channel_float<BIN>(L, c, q, sh, rec_f, C, lab);
  } else {
    unsigned iv[BIN][OA_PX];
#pragma unroll
    for (int r = 0; r < BIN; ++r) unpack_u16x8(m[r], iv[r]);
    if (store_ok) {
      unsigned bs[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) bs[j] = 0u;
#pragma unroll
      for (int r = 0; r < BIN; ++r)
#pragma unroll
        for (int i = 0; i < OA_PX; ++i) bs[i / BIN] += iv[r][i];
      if (NB == 8) {
        stg128_stream(bp, make_uint4(bs[0], bs[1 % NB], bs[2 % NB], bs[3 % NB]), pol_stream);
        stg128_stream(bp + 16, make_uint4(bs[4 % NB], bs[5 % NB], bs[6 % NB], bs[7 % NB]), pol_stream);
      } else if (NB == 4) {
        stg128_stream(bp, make_uint4(bs[0], bs[1 % NB], bs[2 % NB], bs[3 % NB]), pol_stream);
      } else {
        stg64_stream(bp, make_uint2(bs[0], bs[1 % NB]), pol_stream);
      }
    }
    if (warp_fg) f2_channel_int<BIN>(L, c, iv, sh, rec_f, C, lab);
  }
}

// ---- kernel A: direct 128-bit loads ------------------------------------------------------------------
// Grid (F, tiles_x, tiles_y): blockIdx.x = field is the fastest index, so the F blocks that read the
// same piece of the plate-constant illumination function are adjacent in launch order (it comes
// from HBM once per launch and from L2 for the other fields), and no index division is needed.
// Addressing: per-lane byte pointers for the raw stack and the function that advance by a uniform
// stride per channel; rows and z-planes are uniform 32-bit byte offsets from them.  Lanes beyond
// the image width read the last valid column group instead (no default values to materialise),
// carry background labels and skip their stores.
template <int BIN, int ZT, bool HAS_ILLUM>
__global__ void __launch_bounds__(OA_THREADS, BIN == 4 ? 4 : 8)
field_fused2_kernel(const uint16_t* __restrict__ raw, const float* __restrict__ illum, int illum_is_rcp,
                    const void* __restrict__ labels, int label_bytes, uint16_t* __restrict__ maxproj,
                    void* __restrict__ binned, unsigned long long* __restrict__ rec, int* __restrict__ flags,
                    int Nmax, int C, int Z, int H, int W) {
  __shared__ OaShared sh;
  oa_init_shared(sh);
  const int f = blockIdx.x, tile_x = blockIdx.y, tile_y = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = tile_y * OA_WARPS + warp;      // binned row / row block
  const int y0 = rb * BIN;
  const int xw0 = tile_x * 32 * OA_PX;
  const int g = tile_x * 32 + lane;             // 16-byte group along the row
  const int x0 = g * OA_PX;
  const int nz = ZT > 0 ? ZT : Z;
  unsigned long long* rec_f = rec + (size_t)f * Nmax * k3_record_words(C);
  bool overflow = false;
  if (y0 < H) {                                 // warp-uniform (H % BIN == 0: a row block is all in or all out)
    const bool col_ok = x0 < W;
    const unsigned plane = (unsigned)H * (unsigned)W;               // < 2^31 (checked by the caller)
    const unsigned row_b = (unsigned)W * 2u, plane_b = plane * 2u;   // bytes of a uint16 row / plane
    const unsigned off = (unsigned)y0 * (unsigned)W + (unsigned)(col_ok ? x0 : W - OA_PX);
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_keep = policy_evict_last();
    F2Labels lab;
    lab.base = labels; lab.elem = (size_t)f * plane + off; lab.label_bytes = label_bytes; lab.W = W;

    unsigned lw[BIN][F2_WPR];
    f2_load_labels<BIN>(lw, reinterpret_cast<const char*>(labels) + lab.elem * (size_t)label_bytes, label_bytes, row_b,
                        Nmax, col_ok, pol_stream, overflow);
    // walking pointers (bytes), channel 0
    const char* rp = reinterpret_cast<const char*>(raw) + (size_t)f * C * nz * plane_b + (size_t)off * 2;
    const char* ip = reinterpret_cast<const char*>(illum) + (size_t)off * 4;
    constexpr int NB = OA_PX / BIN;
    const unsigned bplane_b = (plane / (BIN * BIN)) * 4u;
    const unsigned boff = ((unsigned)rb * (unsigned)(W / BIN) + (unsigned)g * NB) * 4u;
    const size_t chan_b = (size_t)nz * plane_b;
    const bool pf_raw = (lane & 3) == 0, pf_ill = (lane & 1) == 0;
    // channel 0 towards L2 while the labels are analysed (one lane in four / two touches every line)
    for (int z = 0; z < nz; ++z)
#pragma unroll
      for (int r = 0; r < BIN; ++r)
        if (pf_raw) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (size_t)z * plane_b + r * row_b));
    if (HAS_ILLUM) {
#pragma unroll
      for (int r = 0; r < BIN; ++r)
        if (pf_ill) asm volatile("prefetch.global.L2 [%0];" ::"l"(ip + r * row_b * 2));
    }

    F2Lane<BIN> L;
    L.lane = lane;
    unsigned wany = 0u;
#pragma unroll
    for (int r = 0; r < BIN; ++r)
#pragma unroll
      for (int k = 0; k < F2_WPR; ++k) wany |= lw[r][k];
    const bool warp_fg = __any_sync(OA_FULL, wany != 0u);
    if (warp_fg) f2_begin<BIN>(L, lw, (unsigned)Nmax, y0, xw0, sh, rec_f, C, overflow);

    for (int c = 0; c < C; ++c) {
      uint4 m[BIN];
      uint4 il[BIN][2];
      if (ZT > 0) {
        uint4 v[BIN][ZT > 0 ? ZT : 1];
#pragma unroll
        for (int r = 0; r < BIN; ++r)
#pragma unroll
          for (int z = 0; z < ZT; ++z) v[r][z] = ldg128_stream(rp + (size_t)z * plane_b + r * row_b, pol_stream);
        if (HAS_ILLUM) {
#pragma unroll
          for (int r = 0; r < BIN; ++r) {
            il[r][0] = ldg128_keep(ip + r * row_b * 2, pol_keep);
            il[r][1] = ldg128_keep(ip + r * row_b * 2 + 16, pol_keep);
          }
        }
        if (c + 1 < C) {   // next channel towards L2 (distance 1 measured best)
          const char* np = rp + chan_b;
#pragma unroll
          for (int z = 0; z < ZT; ++z)
#pragma unroll
            for (int r = 0; r < BIN; ++r)
              if (pf_raw) asm volatile("prefetch.global.L2 [%0];" ::"l"(np + (size_t)z * plane_b + r * row_b));
          if (HAS_ILLUM) {
            const char* nip = ip + (size_t)plane_b * 2;
#pragma unroll
            for (int r = 0; r < BIN; ++r)
              if (pf_ill) asm volatile("prefetch.global.L2 [%0];" ::"l"(nip + r * row_b * 2));
          }
        }
#pragma unroll
        for (int r = 0; r < BIN; ++r) {
          m[r] = v[r][0];
#pragma unroll
          for (int z = 1; z < ZT; ++z) m[r] = vmax_u16x8(m[r], v[r][z]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < BIN; ++r) m[r] = ldg128_stream(rp + r * row_b, pol_stream);
        if (HAS_ILLUM) {
#pragma unroll
          for (int r = 0; r < BIN; ++r) {
            il[r][0] = ldg128_keep(ip + r * row_b * 2, pol_keep);
            il[r][1] = ldg128_keep(ip + r * row_b * 2 + 16, pol_keep);
          }
        }
        for (int z = 1; z < nz; ++z) {
#pragma unroll
          for (int r = 0; r < BIN; ++r)
            m[r] = vmax_u16x8(m[r], ldg128_stream(rp + (size_t)z * plane_b + r * row_b, pol_stream));
        }
        if (c + 1 < C) {
          const char* np = rp + chan_b;
          for (int z = 0; z < nz; ++z)
#pragma unroll
            for (int r = 0; r < BIN; ++r)
              if (pf_raw) asm volatile("prefetch.global.L2 [%0];" ::"l"(np + (size_t)z * plane_b + r * row_b));
          if (HAS_ILLUM) {
            const char* nip = ip + (size_t)plane_b * 2;
#pragma unroll
            for (int r = 0; r < BIN; ++r)
              if (pf_ill) asm volatile("prefetch.global.L2 [%0];" ::"l"(nip + r * row_b * 2));
          }
        }
      }
      const size_t fc = (size_t)f * C + c;
      if (maxproj != nullptr && col_ok) {
        char* mp = reinterpret_cast<char*>(maxproj) + fc * plane_b + (size_t)off * 2;
#pragma unroll
        for (int r = 0; r < BIN; ++r) stg128_stream(mp + r * row_b, m[r], pol_stream);
      }
      f2_consume<BIN, HAS_ILLUM>(m, il, illum_is_rcp, binned != nullptr && col_ok,
                                 reinterpret_cast<char*>(binned) + fc * bplane_b + boff, pol_stream, L, warp_fg, c, sh,
                                 rec_f, C, lab);
      rp += chan_b;
      ip += (size_t)plane_b * 2;
    }
  }
  if (overflow) atomicOr(flags + f, 1);
  oa_finish<HAS_ILLUM>(sh, rec_f, C);
}

// ---- kernel B: rows staged in shared memory by the TMA unit ---------------------------------------------
// Same tiling, same arithmetic; the difference is how a channel's rows reach the lanes.  Every warp
// owns one stage in shared memory (Z x BIN raw rows + BIN function rows of its 256-column span).  The
// lanes read a channel's rows out of the stage at the top of the channel (LDS.128, immediate offsets);
// right after that one elected lane hands the rows of channel c + 1 to the TMA unit (cp.async.bulk,
// mbarrier complete_tx), so a whole channel of arithmetic covers the HBM latency of the next one.  No
// registers are tied up by loads in flight (the direct kernel holds 40 of its 64 registers for them),
// no per-lane address arithmetic is needed for inputs, and occupancy no longer has to hide memory
// latency.
constexpr int F2S_CAP = 48;                       // records per CTA (overflowing run heads flush directly)
typedef OaSharedT<F2S_CAP> F2sShared;

__device__ __forceinline__ uint32_t f2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f2_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f2_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void f2_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(f2_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f2_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(f2_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void f2_bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar,
                                             uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(f2_smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ uint4 f2_lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

constexpr int F2S_RAW_ROW = 32 * OA_PX * 2;       // bytes of a warp's span of one uint16 row (512)
constexpr int F2S_ILL_ROW = 32 * OA_PX * 4;       // ... of one float32 row (1024)
__host__ __device__ constexpr size_t f2s_stage_bytes(int bin, int nz, bool has_illum) {
  return (size_t)bin * ((size_t)nz * F2S_RAW_ROW + (has_illum ? F2S_ILL_ROW : 0));
}
__host__ __device__ constexpr size_t f2s_records_bytes() { return (sizeof(F2sShared) + 127) / 128 * 128; }
static size_t f2s_smem_bytes(int bin, int nz, bool has_illum) {
  return f2s_records_bytes() + (size_t)OA_WARPS * f2s_stage_bytes(bin, nz, has_illum) + OA_WARPS * sizeof(uint64_t);
}
__device__ __forceinline__ bool f2_elect_one() {
  uint32_t e;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(e));
  return e != 0u;
}

template <int BIN, int ZT, bool HAS_ILLUM>
__global__ void __launch_bounds__(OA_THREADS, BIN == 4 ? 3 : 7)
field_fused2s_kernel(const uint16_t* __restrict__ raw, const float* __restrict__ illum, int illum_is_rcp,
                     const void* __restrict__ labels, int label_bytes, uint16_t* __restrict__ maxproj,
                     void* __restrict__ binned, unsigned long long* __restrict__ rec, int* __restrict__ flags,
                     int Nmax, int C, int Z, int H, int W) {
  extern __shared__ __align__(128) unsigned char f2s_smem[];
  F2sShared& sh = *reinterpret_cast<F2sShared*>(f2s_smem);
  const int nz = ZT > 0 ? ZT : Z;
  const uint32_t stage_b = (uint32_t)f2s_stage_bytes(BIN, nz, HAS_ILLUM);
  // the shuffle tells the compiler the warp index is warp-uniform: everything the TMA issue needs
  // (stage address, row pointers, byte counts) then lives in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  unsigned char* ring = f2s_smem + f2s_records_bytes() + (size_t)warp * stage_b;
  uint64_t* bar = reinterpret_cast<uint64_t*>(f2s_smem + f2s_records_bytes() + (size_t)OA_WARPS * stage_b) + warp;
  if (lane == 0) {
    f2_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  oa_init_shared(sh);                            // __syncthreads inside
  const int f = blockIdx.x, tile_x = blockIdx.y, tile_y = blockIdx.z;
  const int rb = tile_y * OA_WARPS + warp;
  const int y0 = rb * BIN;
  const int xw0 = tile_x * 32 * OA_PX;
  const int g = tile_x * 32 + lane;
  const int x0 = g * OA_PX;
  unsigned long long* rec_f = rec + (size_t)f * Nmax * k3_record_words(C);
  bool overflow = false;
  if (y0 < H) {
    const bool col_ok = x0 < W;
    const unsigned plane = (unsigned)H * (unsigned)W;
    const unsigned row_b = (unsigned)W * 2u, plane_b = plane * 2u;
    const unsigned woff = (unsigned)y0 * (unsigned)W + (unsigned)xw0;          // the warp's first pixel
    const unsigned off = (unsigned)y0 * (unsigned)W + (unsigned)(col_ok ? x0 : W - OA_PX);
    const uint32_t raw_n = (uint32_t)min(32 * OA_PX, W - xw0) * 2u;             // bytes of the warp's span inside the image
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_keep = policy_evict_last();
    const char* wrp = reinterpret_cast<const char*>(raw) + (size_t)f * C * nz * plane_b + (size_t)woff * 2;
    const char* wip = reinterpret_cast<const char*>(illum) + (size_t)woff * 4;
    const size_t chan_b = (size_t)nz * plane_b;
    const uint32_t ring_s = f2_smem_u32(ring);

    // one elected lane hands the rows of the next channel to the TMA unit; the mbarrier counts the bytes
    const char* nrp = wrp;                       // rows of the next channel to issue (uniform, walks by chan_b)
    const char* nip = wip;
    const uint32_t tx_bytes = (uint32_t)BIN * ((uint32_t)nz * raw_n + (HAS_ILLUM ? 2u * raw_n : 0u));
    auto issue = [&]() {
      if (f2_elect_one()) {
        f2_mbar_expect_tx(bar, tx_bytes);
        const char* rp = nrp;
        uint32_t dst = ring_s;
        for (int z = 0; z < nz; ++z) {
#pragma unroll
          for (int r = 0; r < BIN; ++r) f2_bulk_load(dst + (uint32_t)r * F2S_RAW_ROW, rp + r * row_b, raw_n, bar, pol_stream);
          rp += plane_b;
          dst += (uint32_t)BIN * F2S_RAW_ROW;
        }
        if (HAS_ILLUM) {
#pragma unroll
          for (int r = 0; r < BIN; ++r) f2_bulk_load(dst + (uint32_t)r * F2S_ILL_ROW, nip + r * row_b * 2, 2u * raw_n, bar, pol_keep);
        }
      }
      nrp += chan_b;
      nip += (size_t)plane_b * 2;
    };
    issue();

    F2Labels lab;
    lab.base = labels; lab.elem = (size_t)f * plane + off; lab.label_bytes = label_bytes; lab.W = W;
    unsigned lw[BIN][F2_WPR];
    f2_load_labels<BIN>(lw, reinterpret_cast<const char*>(labels) + lab.elem * (size_t)label_bytes, label_bytes, row_b,
                        Nmax, col_ok, pol_stream, overflow);
    F2Lane<BIN> L;
    L.lane = lane;
    unsigned wany = 0u;
#pragma unroll
    for (int r = 0; r < BIN; ++r)
#pragma unroll
      for (int k = 0; k < F2_WPR; ++k) wany |= lw[r][k];
    const bool warp_fg = __any_sync(OA_FULL, wany != 0u);
    if (warp_fg) f2_begin<BIN>(L, lw, (unsigned)Nmax, y0, xw0, sh, rec_f, C, overflow);

    constexpr int NB = OA_PX / BIN;
    const unsigned bplane_b = (plane / (BIN * BIN)) * 4u;
    const unsigned boff = ((unsigned)rb * (unsigned)(W / BIN) + (unsigned)g * NB) * 4u;
    const uint32_t lane_raw = (uint32_t)lane * 16u, lane_ill = (uint32_t)lane * 32u;
    for (int c = 0; c < C; ++c) {
      f2_mbar_wait(bar, (uint32_t)c & 1u);
      const uint32_t st = ring_s;
      uint4 m[BIN];
      uint4 il[BIN][2];
#pragma unroll
      for (int r = 0; r < BIN; ++r) {
        m[r] = f2_lds128(st + (uint32_t)r * F2S_RAW_ROW + lane_raw);
        if (ZT > 0) {
#pragma unroll
          for (int z = 1; z < ZT; ++z) m[r] = vmax_u16x8(m[r], f2_lds128(st + (uint32_t)(z * BIN + r) * F2S_RAW_ROW + lane_raw));
        } else {
          for (int z = 1; z < nz; ++z) m[r] = vmax_u16x8(m[r], f2_lds128(st + (uint32_t)(z * BIN + r) * F2S_RAW_ROW + lane_raw));
        }
        if (HAS_ILLUM) {
          const uint32_t ia = st + (uint32_t)(nz * BIN) * F2S_RAW_ROW + (uint32_t)r * F2S_ILL_ROW + lane_ill;
          il[r][0] = f2_lds128(ia);
          il[r][1] = f2_lds128(ia + 16u);
        }
      }
      if (c + 1 < C) {
        __syncwarp();                            // every lane has its rows in registers: the stage is free
        issue();
      }
      const size_t fc = (size_t)f * C + c;
      if (maxproj != nullptr && col_ok) {
        char* mp = reinterpret_cast<char*>(maxproj) + fc * plane_b + (size_t)off * 2;
#pragma unroll
        for (int r = 0; r < BIN; ++r) stg128_stream(mp + r * row_b, m[r], pol_stream);
      }
      f2_consume<BIN, HAS_ILLUM>(m, il, illum_is_rcp, binned != nullptr && col_ok,
                                 reinterpret_cast<char*>(binned) + fc * bplane_b + boff, pol_stream, L, warp_fg, c, sh,
                                 rec_f, C, lab);
    }
  }
  if (overflow) atomicOr(flags + f, 1);
  oa_finish<HAS_ILLUM>(sh, rec_f, C);
}

__global__ void __launch_bounds__(256)
illum_reciprocal_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __frcp_rn(in[i]);
}

static bool use_staged_kernel() {
  static const bool v = [] {
    const char* e = getenv("IPS_FUSED_STAGED");      // 1 = TMA-staged kernel, 0 = direct loads (A/B measurements)
    return e == nullptr ? true : e[0] != '0';
  }();
  return v;
}

template <int BIN, bool HAS_ILLUM>
static int launch_fused2(int Z, dim3 grid, cudaStream_t st, const uint16_t* raw, const float* illum, int is_rcp,
                         const void* labels, int label_bytes, uint16_t* maxproj, void* binned,
                         unsigned long long* rec, int* flags, int Nmax, int C, int H, int W) {
  const size_t smem = f2s_smem_bytes(BIN, Z, HAS_ILLUM);
  const bool staged = use_staged_kernel() && smem <= 200 * 1024;
#define IPS_F2_CASE(ZT)                                                                                           \
  do {                                                                                                            \
    if (staged) {                                                                                                 \
      static size_t attr_smem = 0;     /* one device per process (one rank per GPU) */                            \
      if (smem > attr_smem) {                                                                                     \
        IPS_CUDA_OK(cudaFuncSetAttribute(field_fused2s_kernel<BIN, ZT, HAS_ILLUM>,                                \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
        IPS_CUDA_OK(cudaFuncSetAttribute(field_fused2s_kernel<BIN, ZT, HAS_ILLUM>,                                \
                                         cudaFuncAttributePreferredSharedMemoryCarveout, 100));                   \
        attr_smem = smem;                                                                                         \
      }                                                                                                           \
      field_fused2s_kernel<BIN, ZT, HAS_ILLUM><<<grid, OA_THREADS, smem, st>>>(raw, illum, is_rcp, labels, label_bytes, \
                                                                               maxproj, binned, rec, flags, Nmax, C, Z, H, W); \
    } else {                                                                                                      \
      field_fused2_kernel<BIN, ZT, HAS_ILLUM><<<grid, OA_THREADS, 0, st>>>(raw, illum, is_rcp, labels, label_bytes, \
                                                                           maxproj, binned, rec, flags, Nmax, C, Z, H, W); \
    }                                                                                                             \
  } while (0)
  switch (Z) {
    case 3: IPS_F2_CASE(3); break;
    case 5: IPS_F2_CASE(5); break;
    default: IPS_F2_CASE(0); break;
  }
#undef IPS_F2_CASE
  return IPS_OK;
}

// defined in object_stats.cu
int k3_launch_init(unsigned long long* rec, int* flags, int F, int C, int Nmax, cudaStream_t st);
int k3_launch_compact(const unsigned long long* rec, const int* flags, int32_t* n_objects, int32_t* ints,
                      float* flts, int Nmax, int F, int C, float intensity_scale, bool has_illum, cudaStream_t st);
size_t k3_records_bytes_pub(int F, int C, int Nmax);

// Launches the second-generation kernel when the shape allows it; returns 1 if it did, 0 if the
// caller has to take the general path, a negative status on error.
int field_fused2_try(const uint16_t* raw, const float* illum, int illum_is_rcp, const void* labels, int label_bytes,
                     uint16_t* maxproj, void* binned, int bin, float intensity_scale, int32_t* n_objects,
                     int32_t* ints, float* flts, int Nmax, void* ws, int F, int C, int Z, int H, int W,
                     cudaStream_t st) {
  const bool vec = (W % 8 == 0) && aligned16(raw) && aligned16(illum) && aligned16(labels) && aligned16(maxproj) &&
                   (binned == nullptr || (reinterpret_cast<uintptr_t>(binned) & (bin == 4 ? 7u : 15u)) == 0);
  if (!vec || Nmax > 65535 || (size_t)H * W >= (1ull << 31)) return 0;
  unsigned long long* rec = reinterpret_cast<unsigned long long*>(ws);
  int* flags = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + k3_records_bytes_pub(F, C, Nmax));
  int rc = k3_launch_init(rec, flags, F, C, Nmax, st);
  if (rc != IPS_OK) return rc;
  const int tiles_x = (W + 32 * OA_PX - 1) / (32 * OA_PX);
  const int tiles_y = (H / bin + OA_WARPS - 1) / OA_WARPS;
  if (tiles_x > 65535 || tiles_y > 65535) return 0;
  const dim3 grid((unsigned)F, (unsigned)tiles_x, (unsigned)tiles_y);
  const bool has_illum = illum != nullptr;
#define IPS_F2_BIN(B)                                                                                              \
  do {                                                                                                             \
    if (has_illum) rc = launch_fused2<B, true>(Z, grid, st, raw, illum, illum_is_rcp, labels, label_bytes, maxproj, binned, rec, flags, Nmax, C, H, W); \
    else rc = launch_fused2<B, false>(Z, grid, st, raw, illum, 0, labels, label_bytes, maxproj, binned, rec, flags, Nmax, C, H, W); \
  } while (0)
  if (bin == 1) IPS_F2_BIN(1);
  else if (bin == 2) IPS_F2_BIN(2);
  else IPS_F2_BIN(4);
#undef IPS_F2_BIN
  if (rc != IPS_OK) return rc;
  IPS_LAUNCH_OK("field_fused2_kernel");
  rc = k3_launch_compact(rec, flags, n_objects, ints, flts, Nmax, F, C, intensity_scale, has_illum, st);
  return rc == IPS_OK ? 1 : rc;
}

}  // namespace ips

using namespace ips;

extern "C" int ips_illum_reciprocal(const float* illum, float* rcp_out, int64_t n, ips_stream_t stream) {
  if (!illum || !rcp_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_illum_reciprocal: NULL pointer argument");
  if (n < 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_illum_reciprocal: negative size");
  if (n == 0) return IPS_OK;
  illum_reciprocal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(illum, rcp_out, (size_t)n);
  IPS_LAUNCH_OK("illum_reciprocal_kernel");
  return IPS_OK;
}
