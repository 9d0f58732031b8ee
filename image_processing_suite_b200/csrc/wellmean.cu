// Well-level aggregation of per-object rows: df.groupby("Metadata_Well").agg("mean" | "median")
// (Normalize_CP_ami.py:126, Pycyto_pertime.py:69-72) -- the consumer of the all-gather.
//
// Mean.  rows [N][D] with a well id per row -> per-well float64 means, NaN skipped per column as
// pandas does (every (well, column) keeps its own count of non-NaN values).  A CTA takes a tile of
// subs x rpt consecutive rows; a thread owns one column and every subs-th row of the tile, so the
// CTA's loads sweep the tile contiguously and a thread has eight independent loads in flight.
// Rows of one well (the common layout: rows arrive grouped by field, fields by well) fold in
// registers and leave with one atomic per (thread, well, class).  Any number of columns: grid.y
// walks column tiles of at most 256.  Header-led block tables (ips_pack_rows_block) put one table
// block on grid.z: the header is read once per CTA and the CTAs behind the block's count exit.
//
// Float32 rows are summed EXACTLY in 64-bit integers, one per exponent class (wellmean_exact.cuh),
// hence independently of the order in which threads, chunks and ranks deliver them; only the final
// sum over the 32 classes, taken in ascending order by one thread, rounds.  The per-well means of a
// gathered table are therefore bit-identical to those of the local rows (bench.py
// `aggregation.check`), whatever the atomics' order.  +-inf are kept as two flag bits per (well,
// column): the mean is +-inf, or NaN when both occur, as in IEEE addition.
// Float64 rows (the script tables) use one accumulator and ordinary float64 atomics.
//
// Median.  One 256-thread block per (well, 32-column tile) over rows grouped by well through a
// permutation: exact k-th smallest by 8 x 8-bit radix select on order-preserving 64-bit keys,
// both middle ranks, NaN skipped.
#include "ips_common.cuh"
#include "wellmean_exact.cuh"

namespace ips {

constexpr int WM_THREADS = 256;
constexpr int WM_CLASSES = WMX_CLASSES;   // exponent classes of the exact float32 accumulation
constexpr int WM_UNROLL = 8;              // independent loads per thread
constexpr int WM_RPT_SMALL = 32, WM_RPT_LARGE = 64;   // rows per thread (tile = subs x rpt rows)
constexpr int WM_TILE_MAX = 2048;         // rows per tile (their well ids are staged in shared memory)

__global__ void well_zero_kernel(double* sums, int* counts, size_t n_sums, size_t n_counts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_sums) sums[i] = 0.0;
  if (i < n_counts) counts[i] = 0;
}

struct WmOut {
  double* sums;      // [n_wells][D][K] float64, or the same bytes as int64 class sums
  int* colcnt;       // [n_wells][D] non-NaN values
  int* rowcnt;       // [n_wells]
  int* flags;        // [n_wells][D] bit 0: +inf seen, bit 1: -inf seen (exact path only)
};

// What a thread carries for the well it is in: the exact path's class slots (wellmean_exact.cuh) or
// one float64 sum, the rows taken and the NaNs skipped.
struct WmExact {
  WmxSlots s;
  int k = 0;
};
struct WmPlain {
  double acc = 0.0;
  int k = 0, nan = 0;
};

// where a class sum goes when it leaves a thread: one 64-bit integer atomic
struct WmSink {
  unsigned long long* sums;      // the 32 class sums of (well, column)
  __device__ __forceinline__ void operator()(int c, unsigned long long units) const { atomicAdd(sums + c, units); }
};
__device__ __forceinline__ WmSink wm_sink(const WmOut& o, int w, int D, int d) {
  return WmSink{reinterpret_cast<unsigned long long*>(o.sums) + ((size_t)w * D + d) * WM_CLASSES};
}

__device__ __forceinline__ void wm_flush(const WmExact& t, const WmOut& o, int w, int D, int d) {
  wmx_leave(t.s, wm_sink(o, w, D, d));
  if (t.k > t.s.nan) atomicAdd(&o.colcnt[(size_t)w * D + d], t.k - t.s.nan);
  if (t.s.inf) atomicOr(&o.flags[(size_t)w * D + d], t.s.inf);
  if (d == 0) atomicAdd(&o.rowcnt[w], t.k);
}
__device__ __forceinline__ void wm_flush(const WmPlain& t, const WmOut& o, int w, int D, int d) {
  if (t.k > t.nan) {
    atomicAdd(&o.sums[(size_t)w * D + d], t.acc);
    atomicAdd(&o.colcnt[(size_t)w * D + d], t.k - t.nan);
  }
  if (d == 0) atomicAdd(&o.rowcnt[w], t.k);
}

// One batch of WM_UNROLL values of the thread's column, all of well w.  Exact path: the branch-free
// adds for all of them first, then ONE test for the batch; values no slot took (zero, NaN, inf, a new
// class) are seen again by wmx_slow -- class sums do not depend on the order, so deferring them is free,
// and the hot loop has no control flow (an if per value cost eleven register moves per value at the
// merge points).
__device__ __forceinline__ void wm_batch(WmExact& t, const float (&v)[WM_UNROLL], const WmOut& o, int w, int D, int d) {
  unsigned odd = 0u;
#pragma unroll
  for (int j = 0; j < WM_UNROLL; ++j)
    odd |= wmx_fast(t.s, __float_as_uint(v[j]), (double)v[j]) ? (1u << j) : 0u;
  if (odd) {
    const WmSink sink = wm_sink(o, w, D, d);
#pragma unroll
    for (int j = 0; j < WM_UNROLL; ++j)
      if (odd & (1u << j)) wmx_slow(t.s, __float_as_uint(v[j]), (double)v[j], sink);
  }
}
__device__ __forceinline__ void wm_batch(WmPlain& t, const double (&v)[WM_UNROLL], const WmOut&, int, int, int) {
#pragma unroll
  for (int j = 0; j < WM_UNROLL; ++j) {
    if (v[j] == v[j]) t.acc += v[j];
    else ++t.nan;
  }
}
// the slots of a fresh run opened from its first batch (a thread's run is 64 rows: if the first value
// of every class went through the rare branch, most warp batches would have some lane in it)
__device__ __forceinline__ void wm_seed(WmExact& t, const float (&v)[WM_UNROLL]) {
#pragma unroll
  for (int j = 0; j < WM_UNROLL; ++j) wmx_seed(t.s, __float_as_uint(v[j]));
}
__device__ __forceinline__ void wm_seed(WmPlain&, const double (&)[WM_UNROLL]) {}
// one value (tails of a thread's rows, tiles that span wells)
__device__ __forceinline__ void wm_take(WmExact& t, float v, const WmOut& o, int w, int D, int d) {
  if (wmx_fast(t.s, __float_as_uint(v), (double)v)) wmx_slow(t.s, __float_as_uint(v), (double)v, wm_sink(o, w, D, d));
}
__device__ __forceinline__ void wm_take(WmPlain& t, double v, const WmOut&, int, int, int) {
  if (v == v) t.acc += v;
  else ++t.nan;
}

// The rows of one thread: local rows sub, sub + subs, ... of the tile (`mine` of them), column d.
// UNIFORM: every row of the tile belongs to well w0 (checked by the CTA) -- no per-row test.
template <typename T, typename Run, bool UNIFORM>
__device__ __forceinline__ void wm_thread_rows(const T* __restrict__ pv, const int* s_well, int mine, int stride,
                                               int subs, int w0, const WmOut& out, int D, int d, int n_wells) {
  Run run;
  int cur = UNIFORM ? w0 : -2;                              // -2: no well yet
  auto take = [&](T v, int w) {                             // one row, whatever its well
    if (!UNIFORM && w != cur) {
      if (w < 0 || w >= n_wells) return;                    // rows without a well (id out of range) are dropped
      if (cur >= 0) wm_flush(run, out, cur, D, d);
      cur = w;
      run = Run();
    }
    ++run.k;
    wm_take(run, v, out, cur, D, d);
  };
  // batches of WM_UNROLL rows, the next batch's loads issued before the current one is summed
  const T* p = pv;                                          // walks the thread's rows: one add per load
  T nx[WM_UNROLL];
  const int full = mine / WM_UNROLL;
  if (full > 0) {
#pragma unroll
    for (int j = 0; j < WM_UNROLL; ++j, p += stride) nx[j] = *p;
  }
  for (int b = 0; b < full; ++b) {
    T v[WM_UNROLL];
#pragma unroll
    for (int j = 0; j < WM_UNROLL; ++j) v[j] = nx[j];
    if (b + 1 < full) {
#pragma unroll
      for (int j = 0; j < WM_UNROLL; ++j, p += stride) nx[j] = *p;
    }
    if (UNIFORM) {
      if (b == 0) wm_seed(run, v);
      run.k += WM_UNROLL;
      wm_batch(run, v, out, cur, D, d);
    } else {
#pragma unroll
      for (int j = 0; j < WM_UNROLL; ++j) take(v[j], s_well[(b * WM_UNROLL + j) * subs]);
    }
  }
  for (int i = full * WM_UNROLL; i < mine; ++i, p += stride) take(*p, UNIFORM ? w0 : s_well[i * subs]);
  if (cur >= 0) wm_flush(run, out, cur, D, d);
}

// CTA = tile of subs * rpt consecutive rows (at most WM_TILE_MAX); thread (sub, dl) = column
// blockIdx.y * cols + dl.  well != nullptr: one table of N rows with a well id per row.
// well == nullptr (float32 only): table block blockIdx.z of `block_rows` rows, row 0 a header
// carrying the number of valid rows that follow (two 32-bit words), the well id of a row its
// column 0.  The tile's well ids are staged in shared memory once; a tile inside one well (the
// rule: a well is thousands of consecutive rows) runs without per-row tests.
template <typename T, typename Run>
__global__ void __launch_bounds__(WM_THREADS)
well_accumulate_kernel(const T* __restrict__ rows, const int32_t* __restrict__ well, WmOut out, long long N, int D,
                       int n_wells, int cols, int subs, int rpt, long long block_rows) {
  __shared__ int s_well[WM_TILE_MAX];
  const T* seg = rows;
  long long n = N;
  if (well == nullptr) {
    seg += (size_t)blockIdx.z * block_rows * D;
    const uint32_t* h = reinterpret_cast<const uint32_t*>(seg);
    n = (long long)h[0] | ((long long)h[1] << 32);
    if (n > block_rows - 1) n = block_rows - 1;            // a damaged header cannot lead past its block
    seg += D;
  }
  const int tile = subs * rpt;
  const long long R0 = (long long)blockIdx.x * tile;
  if (R0 >= n) return;                                      // the whole CTA: behind the block's count
  const int span = n - R0 < (long long)tile ? (int)(n - R0) : tile;
  const T* tseg = seg + (size_t)R0 * D;
  const int32_t* twell = well != nullptr ? well + R0 : nullptr;
  const int w0 = twell != nullptr ? twell[0] : (int)tseg[0];
  bool same = w0 >= 0 && w0 < n_wells;
  for (int r = threadIdx.x; r < span; r += WM_THREADS) {
    const int w = twell != nullptr ? twell[r] : (int)tseg[(size_t)r * D];
    s_well[r] = w;
    same = same && w == w0;
  }
  const bool uniform = __syncthreads_and(same);
  const int sub = threadIdx.x / cols, dl = threadIdx.x - sub * cols;
  const int d = blockIdx.y * cols + dl;
  if (sub >= subs || d >= D || sub >= span) return;
  const int mine = (span - sub + subs - 1) / subs;
  const T* pv = tseg + (size_t)sub * D + d;
  if (uniform) wm_thread_rows<T, Run, true>(pv, s_well + sub, mine, subs * D, subs, w0, out, D, d, n_wells);
  else wm_thread_rows<T, Run, false>(pv, s_well + sub, mine, subs * D, subs, w0, out, D, d, n_wells);
}

__global__ void well_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ colcnt,
                                     const int* __restrict__ rowcnt, const int* __restrict__ flags,
                                     double* __restrict__ mean_out, int32_t* __restrict__ count_out, int n_wells, int D,
                                     int exact) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_wells * D) return;
  const int w = (int)(i / D);
  const double nan = __longlong_as_double(0x7ff8000000000000ll), inf = __longlong_as_double(0x7ff0000000000000ll);
  double s;
  if (exact) {
    s = wmx_total(reinterpret_cast<const long long*>(sums) + i * WM_CLASSES);
    const int f = flags[i];
    if (f) s = f == 1 ? inf : f == 2 ? -inf : nan;
  } else {
    s = sums[i];
  }
  const int c = colcnt[i];
  mean_out[i] = c > 0 ? s / (double)c : nan;
  if (i % D == 0) count_out[w] = rowcnt[w];
}

static size_t wm_sums_bytes(int n_wells, int D) { return round_up((size_t)n_wells * D * WM_CLASSES * sizeof(double), 256); }
static size_t wm_colcnt_bytes(int n_wells, int D) { return round_up((size_t)n_wells * D * sizeof(int), 256); }
static size_t wm_rowcnt_bytes(int n_wells) { return round_up((size_t)n_wells * sizeof(int), 256); }

// workspace: sums | colcnt | flags | rowcnt (the three integer arrays are adjacent: one zero fill)
static WmOut wm_carve(void* ws, int n_wells, int D) {
  char* p = reinterpret_cast<char*>(ws);
  WmOut w;
  w.sums = reinterpret_cast<double*>(p);
  p += wm_sums_bytes(n_wells, D);
  w.colcnt = reinterpret_cast<int*>(p);
  p += wm_colcnt_bytes(n_wells, D);
  w.flags = reinterpret_cast<int*>(p);
  p += wm_colcnt_bytes(n_wells, D);
  w.rowcnt = reinterpret_cast<int*>(p);
  return w;
}

// ---- median ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long wm_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);       // ascending keys == ascending values
}
__device__ __forceinline__ double wm_unkey(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

constexpr int WMD_COLS = 32, WMD_RG = 8;   // block = 8 row groups x 32 columns

// rows [N][D] float64, perm [N] (row indices grouped by well), offsets [n_wells + 1].
__global__ void __launch_bounds__(WMD_COLS * WMD_RG)
well_median_kernel(const double* __restrict__ rows, const int64_t* __restrict__ perm,
                   const int64_t* __restrict__ offsets, double* __restrict__ out, int32_t* __restrict__ count_out,
                   int D) {
  __shared__ unsigned hist[WMD_COLS][256 + 1];
  __shared__ unsigned long long prefix_s[WMD_COLS];
  __shared__ unsigned rank_s[WMD_COLS], valid_s[WMD_COLS];
  __shared__ double res_s[2][WMD_COLS];
  const int w = blockIdx.x;
  const int cl = threadIdx.x & (WMD_COLS - 1), rg = threadIdx.x / WMD_COLS;
  const int d = blockIdx.y * WMD_COLS + cl;
  const long long b = offsets[w], e = offsets[w + 1];
  const bool col_ok = d < D;
  // non-NaN count per column
  if (rg == 0) valid_s[cl] = 0u;
  __syncthreads();
  {
    unsigned c = 0u;
    if (col_ok)
      for (long long i = b + rg; i < e; i += WMD_RG) {
        const double v = rows[(size_t)perm[i] * D + d];
        c += (v == v);
      }
    if (c) atomicAdd(&valid_s[cl], c);
  }
  __syncthreads();
  const unsigned m = valid_s[cl];
  for (int which = 0; which < 2; ++which) {
    // rank (0-based) of the lower / upper middle value among the m non-NaN ones
    if (rg == 0) { prefix_s[cl] = 0ull; rank_s[cl] = m ? (which == 0 ? (m - 1) / 2 : m / 2) : 0u; }
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      for (int j = rg; j < 256; j += WMD_RG) hist[cl][j] = 0u;
      __syncthreads();
      const unsigned long long pre = prefix_s[cl];
      const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
      if (col_ok && m)
        for (long long i = b + rg; i < e; i += WMD_RG) {
          const double v = rows[(size_t)perm[i] * D + d];
          if (v == v) {
            const unsigned long long k = wm_key(v);
            if ((k & himask) == pre) atomicAdd(&hist[cl][(unsigned)(k >> shift) & 255u], 1u);
          }
        }
      __syncthreads();
      if (rg == 0 && m) {
        unsigned r = rank_s[cl], acc = 0u;
        int dgt = 0;
        for (; dgt < 256; ++dgt) {
          const unsigned h = hist[cl][dgt];
          if (acc + h > r) break;
          acc += h;
        }
        rank_s[cl] = r - acc;
        prefix_s[cl] = pre | ((unsigned long long)dgt << shift);
      }
      __syncthreads();
    }
    if (rg == 0) res_s[which][cl] = m ? wm_unkey(prefix_s[cl]) : __longlong_as_double(0x7ff8000000000000ll);
    __syncthreads();
  }
  if (rg == 0 && col_ok) {
    // pandas / numpy: mean of the two middle values
    out[(size_t)w * D + d] = m ? 0.5 * (res_s[0][cl] + res_s[1][cl]) : res_s[0][cl];
    if (d == 0) count_out[w] = (int32_t)(e - b);
  }
}

}  // namespace ips

using namespace ips;

extern "C" size_t ips_well_mean_workspace_bytes(int n_wells, int D) {
  if (n_wells <= 0 || D <= 0) return 0;
  return wm_sums_bytes(n_wells, D) + 2 * wm_colcnt_bytes(n_wells, D) + wm_rowcnt_bytes(n_wells);
}

static int wm_check(const void* ws, size_t ws_bytes, int D, int n_wells, const char* who) {
  if (ws == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "%s: NULL workspace", who);
  if (D <= 0 || n_wells <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "%s: bad shape D=%d n_wells=%d", who, D, n_wells);
  const size_t need = ips_well_mean_workspace_bytes(n_wells, D);
  if (ws_bytes < need) IPS_FAIL(IPS_ERR_NOMEM, "%s: needs %zu workspace bytes (got %zu)", who, need, ws_bytes);
  if (!aligned16(ws)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "%s: workspace not 16-byte aligned", who);
  return IPS_OK;
}

// The streaming form: reset once, add row blocks as they arrive (e.g. one all-gather chunk at a
// time, overlapping the next chunk's transfer), finalize once.
extern "C" int ips_well_sums_reset(void* ws, size_t ws_bytes, int D, int n_wells, ips_stream_t stream) {
  const int rc = wm_check(ws, ws_bytes, D, n_wells, "ips_well_sums_reset");
  if (rc != IPS_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const WmOut w = wm_carve(ws, n_wells, D);
  const size_t n_sums = (size_t)n_wells * D * WM_CLASSES;
  // colcnt, flags and rowcnt are adjacent up to padding: zero the whole integer tail
  const size_t n_counts = (2 * wm_colcnt_bytes(n_wells, D) + wm_rowcnt_bytes(n_wells)) / sizeof(int);
  well_zero_kernel<<<(unsigned)((n_sums + 255) / 256), 256, 0, st>>>(w.sums, w.colcnt, n_sums, n_counts);
  IPS_LAUNCH_OK("well_zero_kernel");
  return IPS_OK;
}

// n_seg table blocks of seg_rows rows each (header mode, well == nullptr), or one table of N rows.
template <typename T, typename Run>
static int wm_add(const T* rows, const int32_t* well, int64_t N, int64_t n_seg, void* ws, int D, int n_wells,
                  long long block_rows, cudaStream_t st) {
  const WmOut w = wm_carve(ws, n_wells, D);
  const int cols = D < WM_THREADS ? D : WM_THREADS, subs = WM_THREADS / cols;
  const long long seg_rows = well != nullptr ? (long long)N : block_rows - 1;
  int rpt = seg_rows >= (1 << 14) ? WM_RPT_LARGE : WM_RPT_SMALL;
  if (rpt * subs > WM_TILE_MAX) rpt = WM_TILE_MAX / subs;   // narrow tables: subs up to 256
  const long long tile = (long long)subs * rpt;
  const long long blocks = (seg_rows + tile - 1) / tile;
  const int tiles = (D + cols - 1) / cols;
  if (blocks > 0x7fffffffLL || tiles > 65535 || n_seg > 65535)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "well sums: table too large for one call");
  well_accumulate_kernel<T, Run><<<dim3((unsigned)blocks, (unsigned)tiles, (unsigned)n_seg), WM_THREADS, 0, st>>>(
      rows, well, w, (long long)N, D, n_wells, cols, subs, rpt, block_rows);
  IPS_LAUNCH_OK("well_accumulate_kernel");
  return IPS_OK;
}

extern "C" int ips_well_sums_add(const float* rows, const int32_t* well, int64_t N, void* ws, size_t ws_bytes, int D,
                                 int n_wells, ips_stream_t stream) {
  const int rc = wm_check(ws, ws_bytes, D, n_wells, "ips_well_sums_add");
  if (rc != IPS_OK) return rc;
  if (N < 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_sums_add: negative row count");
  if (N == 0) return IPS_OK;
  if (!rows || !well) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_well_sums_add: NULL rows");
  return wm_add<float, WmExact>(rows, well, N, 1, ws, D, n_wells, 0, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ips_well_sums_add_blocks(const float* table, int64_t n_blocks, int64_t block_rows, void* ws,
                                        size_t ws_bytes, int D, int n_wells, ips_stream_t stream) {
  const int rc = wm_check(ws, ws_bytes, D, n_wells, "ips_well_sums_add_blocks");
  if (rc != IPS_OK) return rc;
  if (n_blocks < 0 || block_rows < 1) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_sums_add_blocks: bad block shape");
  if (D < 2) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_sums_add_blocks: header rows need D >= 2");
  if (n_blocks == 0 || block_rows == 1) return IPS_OK;
  if (!table) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_well_sums_add_blocks: NULL table");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // grid.z carries the table blocks, at most 65535 per launch
  for (int64_t b0 = 0; b0 < n_blocks; b0 += 65535) {
    const int64_t nb = n_blocks - b0 < 65535 ? n_blocks - b0 : 65535;
    const int rc2 = wm_add<float, WmExact>(table + (size_t)b0 * block_rows * D, nullptr, 0, nb, ws, D, n_wells, block_rows, st);
    if (rc2 != IPS_OK) return rc2;
  }
  return IPS_OK;
}

static void wm_finalize(const WmOut& w, double* mean_out, int32_t* count_out, int D, int n_wells, int exact,
                        cudaStream_t st) {
  const size_t n = (size_t)n_wells * D;
  well_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.sums, w.colcnt, w.rowcnt, w.flags, mean_out,
                                                                    count_out, n_wells, D, exact);
}

extern "C" int ips_well_sums_finalize(const void* ws, size_t ws_bytes, double* mean_out, int32_t* count_out, int D,
                                      int n_wells, ips_stream_t stream) {
  const int rc = wm_check(ws, ws_bytes, D, n_wells, "ips_well_sums_finalize");
  if (rc != IPS_OK) return rc;
  if (!mean_out || !count_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_well_sums_finalize: NULL output");
  wm_finalize(wm_carve(const_cast<void*>(ws), n_wells, D), mean_out, count_out, D, n_wells, 1,
              reinterpret_cast<cudaStream_t>(stream));
  IPS_LAUNCH_OK("well_finalize_kernel");
  return IPS_OK;
}

extern "C" int ips_well_mean(const float* rows, const int32_t* well, double* mean_out, int32_t* count_out,
                             int N, int D, int n_wells, void* ws, size_t ws_bytes, ips_stream_t stream) {
  if (N < 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_mean: negative row count");
  int rc = ips_well_sums_reset(ws, ws_bytes, D, n_wells, stream);
  if (rc == IPS_OK) rc = ips_well_sums_add(rows, well, N, ws, ws_bytes, D, n_wells, stream);
  if (rc == IPS_OK) rc = ips_well_sums_finalize(ws, ws_bytes, mean_out, count_out, D, n_wells, stream);
  return rc;
}

extern "C" int ips_well_mean_f64(const double* rows, const int32_t* well, double* mean_out, int32_t* count_out,
                                 int64_t N, int D, int n_wells, void* ws, size_t ws_bytes, ips_stream_t stream) {
  if (N < 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_mean_f64: negative row count");
  if (!mean_out || !count_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_well_mean_f64: NULL output");
  int rc = ips_well_sums_reset(ws, ws_bytes, D, n_wells, stream);
  if (rc != IPS_OK) return rc;
  if (N > 0) {
    if (!rows || !well) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_well_mean_f64: NULL rows");
    // one float64 accumulator per (well, column): the first n_wells * D words of the same workspace
    rc = wm_add<double, WmPlain>(rows, well, N, 1, ws, D, n_wells, 0, reinterpret_cast<cudaStream_t>(stream));
    if (rc != IPS_OK) return rc;
  }
  wm_finalize(wm_carve(ws, n_wells, D), mean_out, count_out, D, n_wells, 0, reinterpret_cast<cudaStream_t>(stream));
  IPS_LAUNCH_OK("well_finalize_kernel");
  return IPS_OK;
}

extern "C" int ips_well_median_f64(const double* rows, const int64_t* perm, const int64_t* offsets, double* median_out,
                                   int32_t* count_out, int64_t N, int D, int n_wells, ips_stream_t stream) {
  if (N < 0 || D <= 0 || n_wells <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_median_f64: bad shape");
  if (!perm || !offsets || !median_out || !count_out || (N > 0 && !rows))
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_well_median_f64: NULL pointer argument");
  const int tiles = (D + WMD_COLS - 1) / WMD_COLS;
  if (tiles > 65535) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_median_f64: too many columns");
  well_median_kernel<<<dim3((unsigned)n_wells, (unsigned)tiles), WMD_COLS * WMD_RG, 0,
                       reinterpret_cast<cudaStream_t>(stream)>>>(rows, perm, offsets, median_out, count_out, D);
  IPS_LAUNCH_OK("well_median_kernel");
  return IPS_OK;
}
