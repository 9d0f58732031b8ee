// Well-level aggregation of per-object rows: df.groupby("Metadata_Well").agg("mean" | "median")
// (Normalize_CP_ami.py:126, Pycyto_pertime.py:69-72) -- the consumer of the all-gather.
//
// Mean.  rows [N][D] with a well id per row -> per-well float64 means, NaN skipped per column as
// pandas does (every (well, column) keeps its own count of non-NaN values).  A thread owns one
// column of a 32-row run; consecutive rows of one well (the common layout: rows arrive grouped by
// field, fields by well) fold in a register and leave with one float64 atomic per run.  Any
// number of columns: grid.y walks column tiles of at most 256.
//
// Float32 rows are summed EXACTLY, hence independently of the order in which runs, chunks and
// ranks deliver them: a float32 is an integer multiple of 2^(E-150) below 2^(E-126) (E = biased
// exponent), so values whose exponents share E >> 3 -- a "class" of 8 binades -- add without
// rounding in float64 for up to 2^22 values per (well, column, class).  There is one float64
// accumulator per class (32 of them); only the final sum over the classes, taken in ascending
// order by one thread, rounds.  The per-well means of a gathered table are therefore bit-identical
// to those of the local rows (bench.py `aggregation.check`), whatever the atomics' order.
// Float64 rows (the script tables) use one accumulator and ordinary float64 atomics.
//
// Median.  One 256-thread block per (well, 32-column tile) over rows grouped by well through a
// permutation: exact k-th smallest by 8 x 8-bit radix select on order-preserving 64-bit keys,
// both middle ranks, NaN skipped.
#include "ips_common.cuh"

namespace ips {

constexpr int WM_ROWS = 32;        // rows per thread run
constexpr int WM_THREADS = 256;
constexpr int WM_CLASSES = 32;     // exponent classes of the exact float32 accumulation

__global__ void well_zero_kernel(double* sums, int* counts, size_t n_sums, size_t n_counts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_sums) sums[i] = 0.0;
  if (i < n_counts) counts[i] = 0;
}

// Well id of table row r.  Plain mode: well[r].  Header mode (well == nullptr): the table is a
// sequence of blocks of `block_rows` rows whose first row is a header carrying the number of
// valid rows that follow (two 32-bit words); the well id is column 0 of the row itself.
template <typename T>
__device__ __forceinline__ int wm_row_well(const T* __restrict__ rows, const int32_t* __restrict__ well,
                                           long long r, int D, long long block_rows) {
  if (well != nullptr) return well[r];
  const long long b = r / block_rows, k = r - b * block_rows;
  if (k == 0) return -1;
  const uint32_t* h = reinterpret_cast<const uint32_t*>(rows + (size_t)b * block_rows * D);
  const long long n = (long long)h[0] | ((long long)h[1] << 32);
  return k <= n ? (int)rows[(size_t)r * D] : -1;
}

// Thread (sub, dl): column blockIdx.y * cols + dl of the 32-row run `sub` of the block.  The
// threads of one run read consecutive values of a row (coalesced).
template <typename T, bool EXACT>
__global__ void __launch_bounds__(WM_THREADS)
well_accumulate_kernel(const T* __restrict__ rows, const int32_t* __restrict__ well, double* __restrict__ sums,
                       int* __restrict__ colcnt, int* __restrict__ rowcnt, long long N, int D, int n_wells,
                       int cols, int subs, long long block_rows) {
  constexpr int K = EXACT ? WM_CLASSES : 1;
  const int sub = threadIdx.x / cols, dl = threadIdx.x - sub * cols;
  if (sub >= subs) return;
  const int d = blockIdx.y * cols + dl;
  if (d >= D) return;
  const long long r0 = ((long long)blockIdx.x * subs + sub) * WM_ROWS;
  if (r0 >= N) return;
  const long long r1 = r0 + WM_ROWS < N ? r0 + WM_ROWS : N;
  int cur = -1, cls = 0, cnt = 0, nrow = 0;
  bool pending = false;
  double acc = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const int w = wm_row_well(rows, well, r, D, block_rows);
    if (w < 0 || w >= n_wells) continue;   // rows without a well (id out of range) are dropped
    if (w != cur) {
      if (cur >= 0) {
        if (pending) atomicAdd(&sums[((size_t)cur * D + d) * K + cls], acc);
        if (cnt) atomicAdd(&colcnt[(size_t)cur * D + d], cnt);
        if (d == 0) atomicAdd(&rowcnt[cur], nrow);
      }
      cur = w; acc = 0.0; cnt = 0; nrow = 0; pending = false;
    }
    ++nrow;
    const T v = rows[(size_t)r * D + d];
    if (v == v) {
      if (EXACT) {
        const int k = (int)((__float_as_uint((float)v) >> 26) & 31u);
        if (pending && k != cls) {
          atomicAdd(&sums[((size_t)cur * D + d) * K + cls], acc);
          acc = 0.0;
        }
        cls = k;
      }
      acc += (double)v;
      ++cnt;
      pending = true;
    }
  }
  if (cur >= 0) {
    if (pending) atomicAdd(&sums[((size_t)cur * D + d) * K + cls], acc);
    if (cnt) atomicAdd(&colcnt[(size_t)cur * D + d], cnt);
    if (d == 0) atomicAdd(&rowcnt[cur], nrow);
  }
}

__global__ void well_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ colcnt,
                                     const int* __restrict__ rowcnt, double* __restrict__ mean_out,
                                     int32_t* __restrict__ count_out, int n_wells, int D, int K) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_wells * D) return;
  const int w = (int)(i / D);
  double s = 0.0;
  for (int k = 0; k < K; ++k) s += sums[i * K + k];      // ascending exponent class: small terms first
  const int c = colcnt[i];
  mean_out[i] = c > 0 ? s / (double)c : __longlong_as_double(0x7ff8000000000000ll);
  if (i % D == 0) count_out[w] = rowcnt[w];
}

static size_t wm_sums_bytes(int n_wells, int D) { return round_up((size_t)n_wells * D * WM_CLASSES * sizeof(double), 256); }
static size_t wm_colcnt_bytes(int n_wells, int D) { return round_up((size_t)n_wells * D * sizeof(int), 256); }
static size_t wm_rowcnt_bytes(int n_wells) { return round_up((size_t)n_wells * sizeof(int), 256); }

struct WmWs {
  double* sums;
  int* colcnt;
  int* rowcnt;
};
static WmWs wm_carve(void* ws, int n_wells, int D) {
  char* p = reinterpret_cast<char*>(ws);
  WmWs w;
  w.sums = reinterpret_cast<double*>(p);
  w.colcnt = reinterpret_cast<int*>(p + wm_sums_bytes(n_wells, D));
  w.rowcnt = reinterpret_cast<int*>(p + wm_sums_bytes(n_wells, D) + wm_colcnt_bytes(n_wells, D));
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
  return wm_sums_bytes(n_wells, D) + wm_colcnt_bytes(n_wells, D) + wm_rowcnt_bytes(n_wells);
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
  const WmWs w = wm_carve(ws, n_wells, D);
  const size_t n_sums = (size_t)n_wells * D * WM_CLASSES;
  // colcnt and rowcnt are adjacent up to padding: zero the whole integer tail
  const size_t n_counts = (wm_colcnt_bytes(n_wells, D) + wm_rowcnt_bytes(n_wells)) / sizeof(int);
  well_zero_kernel<<<(unsigned)((n_sums + 255) / 256), 256, 0, st>>>(w.sums, w.colcnt, n_sums, n_counts);
  IPS_LAUNCH_OK("well_zero_kernel");
  return IPS_OK;
}

template <typename T, bool EXACT>
static int wm_add(const T* rows, const int32_t* well, int64_t N, void* ws, int D, int n_wells, long long block_rows,
                  cudaStream_t st) {
  const WmWs w = wm_carve(ws, n_wells, D);
  const int cols = D < WM_THREADS ? D : WM_THREADS, subs = WM_THREADS / cols;
  const long long rows_per_block = (long long)subs * WM_ROWS;
  const long long blocks = (N + rows_per_block - 1) / rows_per_block;
  const int tiles = (D + cols - 1) / cols;
  if (blocks > 0x7fffffffLL || tiles > 65535) IPS_FAIL(IPS_ERR_BAD_SHAPE, "well sums: table too large for one call");
  well_accumulate_kernel<T, EXACT><<<dim3((unsigned)blocks, (unsigned)tiles), WM_THREADS, 0, st>>>(
      rows, well, w.sums, w.colcnt, w.rowcnt, N, D, n_wells, cols, subs, block_rows);
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
  return wm_add<float, true>(rows, well, N, ws, D, n_wells, 0, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ips_well_sums_add_blocks(const float* table, int64_t n_blocks, int64_t block_rows, void* ws,
                                        size_t ws_bytes, int D, int n_wells, ips_stream_t stream) {
  const int rc = wm_check(ws, ws_bytes, D, n_wells, "ips_well_sums_add_blocks");
  if (rc != IPS_OK) return rc;
  if (n_blocks < 0 || block_rows < 1) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_sums_add_blocks: bad block shape");
  if (D < 2) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_sums_add_blocks: header rows need D >= 2");
  if (n_blocks == 0 || block_rows == 1) return IPS_OK;
  if (!table) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_well_sums_add_blocks: NULL table");
  return wm_add<float, true>(table, nullptr, n_blocks * block_rows, ws, D, n_wells, block_rows,
                             reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ips_well_sums_finalize(const void* ws, size_t ws_bytes, double* mean_out, int32_t* count_out, int D,
                                      int n_wells, ips_stream_t stream) {
  const int rc = wm_check(ws, ws_bytes, D, n_wells, "ips_well_sums_finalize");
  if (rc != IPS_OK) return rc;
  if (!mean_out || !count_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_well_sums_finalize: NULL output");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const WmWs w = wm_carve(const_cast<void*>(ws), n_wells, D);
  const size_t n = (size_t)n_wells * D;
  well_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.sums, w.colcnt, w.rowcnt, mean_out, count_out,
                                                                    n_wells, D, WM_CLASSES);
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
  int rc = ips_well_sums_reset(ws, ws_bytes, D, n_wells, stream);
  if (rc != IPS_OK) return rc;
  if (N > 0) {
    if (!rows || !well) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_well_mean_f64: NULL rows");
    // one accumulator per (well, column): class 0 of the same layout, the other classes stay zero
    const WmWs w = wm_carve(ws, n_wells, D);
    const int cols = D < WM_THREADS ? D : WM_THREADS, subs = WM_THREADS / cols;
    const long long rows_per_block = (long long)subs * WM_ROWS;
    const long long blocks = (N + rows_per_block - 1) / rows_per_block;
    const int tiles = (D + cols - 1) / cols;
    if (blocks > 0x7fffffffLL || tiles > 65535) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_mean_f64: table too large for one call");
    well_accumulate_kernel<double, false><<<dim3((unsigned)blocks, (unsigned)tiles), WM_THREADS, 0,
                                            reinterpret_cast<cudaStream_t>(stream)>>>(
        rows, well, w.sums, w.colcnt, w.rowcnt, N, D, n_wells, cols, subs, 0);
    IPS_LAUNCH_OK("well_accumulate_kernel");
    well_finalize_kernel<<<(unsigned)(((size_t)n_wells * D + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        w.sums, w.colcnt, w.rowcnt, mean_out, count_out, n_wells, D, 1);
    IPS_LAUNCH_OK("well_finalize_kernel");
    return IPS_OK;
  }
  return ips_well_sums_finalize(ws, ws_bytes, mean_out, count_out, D, n_wells, stream);
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
