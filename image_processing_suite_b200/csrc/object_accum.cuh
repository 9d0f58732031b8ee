// Label-keyed segmented reduction machinery shared by K3 (object_stats.cu) and the fused
// field kernel (field_fused.cu).
//
// Decomposition.  A CTA of OA_WARPS (4) warps owns a tile of 256 columns x (4 * ROWS) rows: warp w
// owns ROWS consecutive rows, lane l owns 8 consecutive columns of them (one 128-bit word per
// row and plane).  Cellpose objects are compact, so
//   * within a lane's ROWS x 8 window there are almost always at most two labels: the window
//     is split into slot 1 (first label met), slot 2 (second label) and a rare remainder that
//     is flushed pixel by pixel;
//   * lanes next to each other mostly carry the same slot-1 label: runs of equal labels along
//     the warp are reduced with a segmented shuffle tree (runs are cut every 8 lanes: 3 steps),
//     so one lane per run -- the run head -- holds the run's partial;
//   * the warps of the CTA see the same objects in consecutive rows: run heads append their
//     partials to a record list in shared memory, after a barrier records of equal label are
//     merged, and each distinct label of the tile is flushed to its global accumulator record
//     once (integer atomics for area / bbox / coordinate sums: exact and order independent;
//     64-bit atomics for the intensity moments).
// Intensity moments: a lane accumulates sum(q - p) and sum((q - p)^2) in fp32 around a pivot p
// taken from the slot itself (no cancellation), converts to float64 sum(q), sum(q^2) once per
// slot, and everything downstream (tree, records, global accumulators) is float64.  Without an
// illumination function the moments are exact integers (uint64).
#pragma once
#include <limits.h>

#include "ips_common.cuh"

namespace ips {

constexpr int OA_WARPS = 4;      // measured: 4-warp CTAs at 8 CTAs/SM beat 8-warp CTAs at 4 (shorter barrier waits)
constexpr int OA_THREADS = OA_WARPS * 32;
constexpr int OA_CAP = 96;       // records per CTA; overflowing run heads flush directly
constexpr int OA_PX = 8;
constexpr int OA_CMAX = 8;
constexpr unsigned OA_FULL = 0xffffffffu;
constexpr int OA_SEG = 8;        // a run of equal labels is cut every OA_SEG lanes: 3 tree steps

// global accumulator record, 8-byte words: [0] sum_y, [1] sum_x, [2] area | ymin,
// [3] xmin | ymax1, [4] xmax1 | pad, then per channel [5+3c] sum, [6+3c] sumsq, [7+3c] min | max.
__host__ __device__ constexpr int k3_record_words(int C) { return 5 + 3 * C; }

struct OaGeom {
  unsigned area, sy, sx;
  int ymin, ymax1, xmin, xmax1;
};

template <int CAP_>
struct OaSharedT {
  static constexpr int CAP = CAP_;
  int count;
  int n_owner;
  int label[CAP_];
  int owner[CAP_];
  OaGeom geom[CAP_];
  unsigned long long s1[CAP_][OA_CMAX], s2[CAP_][OA_CMAX];   // float64 bits or uint64
  unsigned mn[CAP_][OA_CMAX], mx[CAP_][OA_CMAX];             // float bits or integers
};
typedef OaSharedT<OA_CAP> OaShared;

// ---- global flushes -----------------------------------------------------------------------
__device__ __forceinline__ void oa_flush_geom(unsigned long long* __restrict__ rec_f, int C, int label,
                                              const OaGeom& g) {
  unsigned long long* r = rec_f + (size_t)(label - 1) * k3_record_words(C);
  unsigned* r32 = reinterpret_cast<unsigned*>(r);
  atomicAdd(r + 0, (unsigned long long)g.sy);
  atomicAdd(r + 1, (unsigned long long)g.sx);
  atomicAdd(r32 + 4, g.area);
  atomicMin(reinterpret_cast<int*>(r32 + 5), g.ymin);
  atomicMin(reinterpret_cast<int*>(r32 + 6), g.xmin);
  atomicMax(reinterpret_cast<int*>(r32 + 7), g.ymax1);
  atomicMax(reinterpret_cast<int*>(r32 + 8), g.xmax1);
}

template <bool FLOAT_MODE>
__device__ __forceinline__ void oa_flush_chan(unsigned long long* __restrict__ rec_f, int C, int label, int c,
                                              unsigned long long s1, unsigned long long s2, unsigned mn,
                                              unsigned mx) {
  unsigned long long* r = rec_f + (size_t)(label - 1) * k3_record_words(C) + 5 + 3 * c;
  if (FLOAT_MODE) {
    atomicAdd(reinterpret_cast<double*>(r), __longlong_as_double((long long)s1));
    atomicAdd(reinterpret_cast<double*>(r + 1), __longlong_as_double((long long)s2));
  } else {
    atomicAdd(r, s1);
    atomicAdd(r + 1, s2);
  }
  unsigned* r32 = reinterpret_cast<unsigned*>(r + 2);
  atomicMin(r32, mn);
  atomicMax(r32 + 1, mx);
}

// ---- per-lane label analysis ----------------------------------------------------------------
// lab[r][i]: labels of the lane's window.  Slot 1 = first label met in scan order, slot 2 = the
// second, m3 = pixels of any further label.  Labels <= 0 are background; labels > Nmax raise
// the overflow flag and are skipped.
template <int ROWS>
__device__ __forceinline__ void oa_analyze(const int (&lab)[ROWS][OA_PX], int Nmax, int& L1, unsigned& m1,
                                           int& L2, unsigned& m2, unsigned& m3, bool& overflow) {
  L1 = 0; L2 = 0; m1 = 0u; m2 = 0u; m3 = 0u;
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int i = 0; i < OA_PX; ++i) {
      const int l = lab[r][i];
      const unsigned bit = 1u << (r * OA_PX + i);
      if (l > Nmax) overflow = true;
      else if (l > 0) {
        if (L1 == 0) L1 = l;
        if (l == L1) m1 |= bit;
        else {
          if (L2 == 0) L2 = l;
          if (l == L2) m2 |= bit;
          else m3 |= bit;
        }
      }
    }
}

template <int ROWS>
__device__ __forceinline__ OaGeom oa_geom(unsigned m, int y0, int x0) {
  OaGeom g;
  g.area = __popc(m);
  g.sy = 0u; g.sx = 0u; g.ymin = INT_MAX; g.ymax1 = 0;
  unsigned cols = 0u;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const unsigned mr = (m >> (r * OA_PX)) & 0xffu;
    const unsigned cnt = __popc(mr);
    cols |= mr;
    g.sy += (unsigned)(y0 + r) * cnt;
    g.sx += (unsigned)x0 * cnt + __popc(mr & 0xAAu) + 2u * __popc(mr & 0xCCu) + 4u * __popc(mr & 0xF0u);
    if (mr) {
      g.ymin = min(g.ymin, y0 + r);
      g.ymax1 = y0 + r + 1;
    }
  }
  g.xmin = cols ? x0 + __ffs(cols) - 1 : INT_MAX;
  g.xmax1 = cols ? x0 + 32 - __clz(cols) : 0;
  return g;
}

// ---- segmented (run of equal slot-1 labels along the warp) shuffle reductions ------------
// Runs are cut at every OA_SEG-th lane, so OA_SEG.log2 shuffle steps reduce any run; a longer
// run simply yields several records of the same label, which oa_finish merges.
struct OaSeg {
  int lane, seg_last;
  bool head;
};

__device__ __forceinline__ OaSeg oa_segments(int L1) {
  OaSeg s;
  s.lane = threadIdx.x & 31;
  const int prev = __shfl_up_sync(OA_FULL, L1, 1);
  s.head = ((s.lane & (OA_SEG - 1)) == 0) || (prev != L1);
  const unsigned heads = __ballot_sync(OA_FULL, s.head);
  const unsigned above = heads & ~((2u << s.lane) - 1u);
  s.seg_last = above ? (__ffs(above) - 2) : 31;
  return s;
}

template <typename T>
__device__ __forceinline__ T oa_seg_add(T v, const OaSeg& s) {
#pragma unroll
  for (int d = 1; d < OA_SEG; d <<= 1) {
    const T o = __shfl_down_sync(OA_FULL, v, d);
    if (s.lane + d <= s.seg_last) v += o;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T oa_seg_min(T v, const OaSeg& s) {
#pragma unroll
  for (int d = 1; d < OA_SEG; d <<= 1) {
    const T o = __shfl_down_sync(OA_FULL, v, d);
    if (s.lane + d <= s.seg_last) v = min(v, o);
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T oa_seg_max(T v, const OaSeg& s) {
#pragma unroll
  for (int d = 1; d < OA_SEG; d <<= 1) {
    const T o = __shfl_down_sync(OA_FULL, v, d);
    if (s.lane + d <= s.seg_last) v = max(v, o);
  }
  return v;
}

__device__ __forceinline__ OaGeom oa_seg_geom(OaGeom g, const OaSeg& s) {
  g.area = oa_seg_add(g.area, s);
  g.sy = oa_seg_add(g.sy, s);
  g.sx = oa_seg_add(g.sx, s);
  g.ymin = oa_seg_min(g.ymin, s);
  g.xmin = oa_seg_min(g.xmin, s);
  g.ymax1 = oa_seg_max(g.ymax1, s);
  g.xmax1 = oa_seg_max(g.xmax1, s);
  return g;
}

// ---- the per-lane state that lives across the channel loop ------------------------------------
template <int ROWS>
struct OaLane {
  int L1, L2;
  unsigned m1, m2, m3;
  int rec1, rec2;          // record index in shared memory, -1: flush directly
  bool lead1, has2;
  OaSeg seg;
  float nw1[ROWS][OA_PX];  // 0 for slot-1 pixels, 1 elsewhere (mask as an FMA operand)
  const int32_t* lab_ptr;  // the lane's first label (re-read on the rare m3 path)
  int lab_stride;
};

// Call once per lane after the labels are loaded (all 32 lanes of the warp must call).
template <int ROWS>
__device__ __forceinline__ void oa_begin(OaLane<ROWS>& L, const int (&lab)[ROWS][OA_PX], const int32_t* lab_ptr,
                                         int lab_stride, int Nmax, int y0, int x0, OaShared& sh,
                                         unsigned long long* __restrict__ rec_f, int C, bool& overflow) {
  oa_analyze<ROWS>(lab, Nmax, L.L1, L.m1, L.L2, L.m2, L.m3, overflow);
  L.lab_ptr = lab_ptr;
  L.lab_stride = lab_stride;
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int i = 0; i < OA_PX; ++i) L.nw1[r][i] = ((L.m1 >> (r * OA_PX + i)) & 1u) ? 0.f : 1.f;
  L.seg = oa_segments(L.L1);
  L.lead1 = L.seg.head && L.L1 != 0;
  L.has2 = L.L2 != 0;
  const unsigned need1 = __ballot_sync(OA_FULL, L.lead1);
  const unsigned need2 = __ballot_sync(OA_FULL, L.has2);
  const int n1 = __popc(need1), n2 = __popc(need2);
  int base = 0;
  if (L.seg.lane == 0 && n1 + n2 > 0) base = atomicAdd(&sh.count, n1 + n2);
  base = __shfl_sync(OA_FULL, base, 0);
  const unsigned lt = (1u << L.seg.lane) - 1u;
  L.rec1 = base + __popc(need1 & lt);
  L.rec2 = base + n1 + __popc(need2 & lt);
  if (!L.lead1 || L.rec1 >= OA_CAP) L.rec1 = -1;
  if (!L.has2 || L.rec2 >= OA_CAP) L.rec2 = -1;

  // geometry: slot 1 through the tree, slot 2 and the remainder directly
  if (__any_sync(OA_FULL, L.L1 != 0)) {
    OaGeom g1 = oa_seg_geom(oa_geom<ROWS>(L.m1, y0, x0), L.seg);
    if (L.lead1) {
      if (L.rec1 >= 0) { sh.label[L.rec1] = L.L1; sh.geom[L.rec1] = g1; }
      else oa_flush_geom(rec_f, C, L.L1, g1);
    }
  }
  if (L.has2) {
    const OaGeom g2 = oa_geom<ROWS>(L.m2, y0, x0);
    if (L.rec2 >= 0) { sh.label[L.rec2] = L.L2; sh.geom[L.rec2] = g2; }
    else oa_flush_geom(rec_f, C, L.L2, g2);
  }
  if (L.m3) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int i = 0; i < OA_PX; ++i)
        if ((L.m3 >> (r * OA_PX + i)) & 1u)
          oa_flush_geom(rec_f, C, lab[r][i], oa_geom<ROWS>(1u << (r * OA_PX + i), y0, x0));
  }
}

// float64 sum(q), sum(q^2) from the fp32 partials around pivot p
__device__ __forceinline__ void oa_to_double(unsigned m, float p, float a1, float a2, unsigned long long& s1,
                                             unsigned long long& s2) {
  const double n = (double)__popc(m), P = (double)p, A1 = (double)a1;
  s1 = (unsigned long long)__double_as_longlong(fma(n, P, A1));
  s2 = (unsigned long long)__double_as_longlong(fma(n * P, P, fma(2.0 * P, A1, (double)a2)));
}

// Slot-1 partial of one channel, float mode.  The slot mask enters as the FMA operand nw
// (0 inside the slot, 1 outside): no predicates, no selects.  Pass A finds min / max, pass B
// accumulates sum(q - p), sum((q - p)^2) around the pivot p = slot minimum.
template <int ROWS>
__device__ __forceinline__ void oa_slot1_float(unsigned m, const float (&nw)[ROWS][OA_PX],
                                               const float (&fv)[ROWS][OA_PX], unsigned long long& s1,
                                               unsigned long long& s2, unsigned& mn, unsigned& mx) {
  const float BIG = 3.0e38f;
  float lo = BIG, hi = 0.f;
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int i = 0; i < OA_PX; ++i) {
      const float q = fv[r][i];
      lo = fminf(lo, fmaf(nw[r][i], BIG, q));
      hi = fmaxf(hi, fmaf(-q, nw[r][i], q));
    }
  const float p = m ? lo : 0.f;
  float a1 = 0.f, a2 = 0.f;
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int i = 0; i < OA_PX; ++i) {
      const float d = fv[r][i] - p;
      const float dm = fmaf(-d, nw[r][i], d);
      a1 += dm;
      a2 = fmaf(dm, d, a2);
    }
  oa_to_double(m, p, a1, a2, s1, s2);
  mn = __float_as_uint(lo);
  mx = __float_as_uint(hi);
}

// Generic slot partial (bit-tested mask): slot 2 in float mode, both slots in integer mode.
template <int ROWS, bool FLOAT_MODE>
__device__ __forceinline__ void oa_slot_partial(unsigned m, const float (&fv)[ROWS][OA_PX],
                                                const unsigned (&iv)[ROWS][OA_PX], unsigned long long& s1,
                                                unsigned long long& s2, unsigned& mn, unsigned& mx) {
  if (FLOAT_MODE) {
    float lo = __int_as_float(0x7f800000), hi = 0.f;
#pragma unroll
    for (int j = 0; j < ROWS * OA_PX; ++j) {
      const bool in = (m >> j) & 1u;
      const float q = fv[j / OA_PX][j % OA_PX];
      lo = fminf(lo, in ? q : lo);
      hi = fmaxf(hi, in ? q : hi);
    }
    const float p = m ? lo : 0.f;
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int j = 0; j < ROWS * OA_PX; ++j) {
      const bool in = (m >> j) & 1u;
      const float d = in ? fv[j / OA_PX][j % OA_PX] - p : 0.f;
      a1 += d;
      a2 = fmaf(d, d, a2);
    }
    oa_to_double(m, p, a1, a2, s1, s2);
    mn = __float_as_uint(lo);
    mx = __float_as_uint(hi);
  } else {
    unsigned a1 = 0u, lo = 0xffffffffu, hi = 0u;
    unsigned long long a2 = 0ull;
#pragma unroll
    for (int j = 0; j < ROWS * OA_PX; ++j) {
      const bool in = (m >> j) & 1u;
      const unsigned x = iv[j / OA_PX][j % OA_PX];
      const unsigned xm = in ? x : 0u;
      a1 += xm;
      a2 += (unsigned long long)(xm * xm);
      lo = min(lo, in ? x : lo);
      hi = max(hi, in ? x : hi);
    }
    s1 = a1; s2 = a2; mn = lo; mx = hi;
  }
}

// Fold channel c of the lane's window into the tile's records (all 32 lanes must call).
template <int ROWS, bool FLOAT_MODE>
__device__ __forceinline__ void oa_channel(const OaLane<ROWS>& L, int c, const float (&fv)[ROWS][OA_PX],
                                           const unsigned (&iv)[ROWS][OA_PX], OaShared& sh,
                                           unsigned long long* __restrict__ rec_f, int C) {
  if (__any_sync(OA_FULL, L.L1 != 0)) {
    unsigned long long s1, s2;
    unsigned mn, mx;
    if (FLOAT_MODE) oa_slot1_float<ROWS>(L.m1, L.nw1, fv, s1, s2, mn, mx);
    else oa_slot_partial<ROWS, false>(L.m1, fv, iv, s1, s2, mn, mx);
    if (FLOAT_MODE) {
      double d1 = oa_seg_add(__longlong_as_double((long long)s1), L.seg);
      double d2 = oa_seg_add(__longlong_as_double((long long)s2), L.seg);
      s1 = (unsigned long long)__double_as_longlong(d1);
      s2 = (unsigned long long)__double_as_longlong(d2);
    } else {
      s1 = oa_seg_add(s1, L.seg);
      s2 = oa_seg_add(s2, L.seg);
    }
    mn = oa_seg_min(mn, L.seg);
    mx = oa_seg_max(mx, L.seg);
    if (L.lead1) {
      if (L.rec1 >= 0) { sh.s1[L.rec1][c] = s1; sh.s2[L.rec1][c] = s2; sh.mn[L.rec1][c] = mn; sh.mx[L.rec1][c] = mx; }
      else oa_flush_chan<FLOAT_MODE>(rec_f, C, L.L1, c, s1, s2, mn, mx);
    }
  }
  if (__any_sync(OA_FULL, L.has2)) {
    if (L.has2) {
      unsigned long long s1, s2;
      unsigned mn, mx;
      oa_slot_partial<ROWS, FLOAT_MODE>(L.m2, fv, iv, s1, s2, mn, mx);
      if (L.rec2 >= 0) { sh.s1[L.rec2][c] = s1; sh.s2[L.rec2][c] = s2; sh.mn[L.rec2][c] = mn; sh.mx[L.rec2][c] = mx; }
      else oa_flush_chan<FLOAT_MODE>(rec_f, C, L.L2, c, s1, s2, mn, mx);
    }
  }
  if (L.m3) {
#pragma unroll
    for (int j = 0; j < ROWS * OA_PX; ++j)
      if ((L.m3 >> j) & 1u) {
        unsigned long long s1, s2;
        unsigned v;
        if (FLOAT_MODE) {
          const float q = fv[j / OA_PX][j % OA_PX];
          const double qd = (double)q;
          s1 = (unsigned long long)__double_as_longlong(qd);
          s2 = (unsigned long long)__double_as_longlong(qd * qd);
          v = __float_as_uint(q);
        } else {
          const unsigned x = iv[j / OA_PX][j % OA_PX];
          s1 = x; s2 = (unsigned long long)(x * x); v = x;
        }
        oa_flush_chan<FLOAT_MODE>(rec_f, C, L.lab_ptr[(j / OA_PX) * L.lab_stride + (j % OA_PX)], c, s1, s2, v, v);
      }
  }
}

template <typename SH>
__device__ __forceinline__ void oa_init_shared(SH& sh) {
  if (threadIdx.x == 0) { sh.count = 0; sh.n_owner = 0; }
  __syncthreads();
}

// After every lane of the CTA has folded all channels: merge records of equal label and flush
// each distinct label of the tile once.
template <bool FLOAT_MODE, typename SH>
__device__ __forceinline__ void oa_finish(SH& sh, unsigned long long* __restrict__ rec_f, int C) {
  __syncthreads();
  const int n = min(sh.count, SH::CAP);
  for (int r = threadIdx.x; r < n; r += OA_THREADS) {
    const int lab = sh.label[r];
    bool own = true;
    for (int k = 0; k < r; ++k)
      if (sh.label[k] == lab) { own = false; break; }
    if (own) sh.owner[atomicAdd(&sh.n_owner, 1)] = r;
  }
  __syncthreads();
  const int n_own = sh.n_owner;
  for (int t = threadIdx.x; t < n_own * (C + 1); t += OA_THREADS) {
    const int r = sh.owner[t / (C + 1)];
    const int part = t % (C + 1);
    const int lab = sh.label[r];
    if (part == 0) {
      OaGeom g = sh.geom[r];
      for (int k = r + 1; k < n; ++k)
        if (sh.label[k] == lab) {
          const OaGeom o = sh.geom[k];
          g.area += o.area; g.sy += o.sy; g.sx += o.sx;
          g.ymin = min(g.ymin, o.ymin); g.xmin = min(g.xmin, o.xmin);
          g.ymax1 = max(g.ymax1, o.ymax1); g.xmax1 = max(g.xmax1, o.xmax1);
        }
      oa_flush_geom(rec_f, C, lab, g);
    } else {
      const int c = part - 1;
      unsigned long long s1 = sh.s1[r][c], s2 = sh.s2[r][c];
      unsigned mn = sh.mn[r][c], mx = sh.mx[r][c];
      for (int k = r + 1; k < n; ++k)
        if (sh.label[k] == lab) {
          if (FLOAT_MODE) {
            s1 = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)s1) +
                                                          __longlong_as_double((long long)sh.s1[k][c]));
            s2 = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)s2) +
                                                          __longlong_as_double((long long)sh.s2[k][c]));
          } else {
            s1 += sh.s1[k][c];
            s2 += sh.s2[k][c];
          }
          mn = min(mn, sh.mn[k][c]);
          mx = max(mx, sh.mx[k][c]);
        }
      oa_flush_chan<FLOAT_MODE>(rec_f, C, lab, c, s1, s2, mn, mx);
    }
  }
}

}  // namespace ips
