// Well-level aggregation of per-object rows: df.groupby("Metadata_Well").agg("mean")
// (Normalize_CP_ami.py:126, Pycyto_pertime.py:69-72) -- the consumer of the all-gather.
//
// rows [N][D] float32 with a well id per row -> per-well float64 means.  A thread owns one
// feature column of a 32-row run; consecutive rows of one well (the common layout: rows
// arrive grouped by field, fields by well) are folded in a register and flushed with one
// float64 atomic per (well, column) run.
#include "ips_common.cuh"

namespace ips {

constexpr int WM_ROWS = 32;        // rows per thread run
constexpr int WM_THREADS = 256;

__global__ void well_zero_kernel(double* sums, int* counts, size_t n_sums, int n_wells) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_sums) sums[i] = 0.0;
  if (i < (size_t)n_wells) counts[i] = 0;
}

// Thread (sub, d): column d (d == D counts rows) of the 32-row run `sub` of the block.  The
// threads of one run read consecutive floats of a row (coalesced); runs of one well fold in
// a register and flush with one atomic.
__global__ void __launch_bounds__(WM_THREADS)
well_accumulate_kernel(const float* __restrict__ rows, const int32_t* __restrict__ well,
                       double* __restrict__ sums, int* __restrict__ counts, long long N, int D, int n_wells,
                       int cols, int subs) {
  const int sub = threadIdx.x / cols, d = threadIdx.x - sub * cols;
  if (sub >= subs) return;
  const long long r0 = ((long long)blockIdx.x * subs + sub) * WM_ROWS;
  if (r0 >= N) return;
  const long long r1 = r0 + WM_ROWS < N ? r0 + WM_ROWS : N;
  int cur = -1, cnt = 0;
  double acc = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const int w = well[r];
    if (w < 0 || w >= n_wells) continue;   // rows without a well (id out of range) are dropped
    if (w != cur) {
      if (cur >= 0) {
        if (d < D) atomicAdd(&sums[(size_t)cur * D + d], acc);
        else atomicAdd(&counts[cur], cnt);
      }
      cur = w;
      acc = 0.0;
      cnt = 0;
    }
    if (d < D) acc += (double)rows[(size_t)r * D + d];
    else ++cnt;
  }
  if (cur >= 0) {
    if (d < D) atomicAdd(&sums[(size_t)cur * D + d], acc);
    else atomicAdd(&counts[cur], cnt);
  }
}

__global__ void well_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ counts,
                                     double* __restrict__ mean_out, int32_t* __restrict__ count_out,
                                     int n_wells, int D) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_wells * D) return;
  const int w = (int)(i / D);
  const int c = counts[w];
  mean_out[i] = c > 0 ? sums[i] / (double)c : __longlong_as_double(0x7ff8000000000000ll);
  if (i % D == 0) count_out[w] = c;
}

static size_t wm_sums_bytes(int n_wells, int D) { return round_up((size_t)n_wells * D * sizeof(double), 256); }
static size_t wm_counts_bytes(int n_wells) { return round_up((size_t)n_wells * sizeof(int), 256); }

}  // namespace ips

using namespace ips;

extern "C" size_t ips_well_mean_workspace_bytes(int n_wells, int D) {
  if (n_wells <= 0 || D <= 0) return 0;
  return wm_sums_bytes(n_wells, D) + wm_counts_bytes(n_wells);
}

static int wm_check(const void* ws, size_t ws_bytes, int D, int n_wells, const char* who) {
  if (ws == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "%s: NULL workspace", who);
  if (D <= 0 || n_wells <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "%s: bad shape D=%d n_wells=%d", who, D, n_wells);
  if (D + 1 > WM_THREADS) IPS_FAIL(IPS_ERR_BAD_SHAPE, "%s: at most %d feature columns (got %d)", who, WM_THREADS - 1, D);
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
  double* sums = reinterpret_cast<double*>(ws);
  int* counts = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + wm_sums_bytes(n_wells, D));
  const size_t n_sums = (size_t)n_wells * D;
  well_zero_kernel<<<(unsigned)((n_sums + 255) / 256), 256, 0, st>>>(sums, counts, n_sums, n_wells);
  IPS_LAUNCH_OK("well_zero_kernel");
  return IPS_OK;
}

extern "C" int ips_well_sums_add(const float* rows, const int32_t* well, int64_t N, void* ws, size_t ws_bytes, int D,
                                 int n_wells, ips_stream_t stream) {
  const int rc = wm_check(ws, ws_bytes, D, n_wells, "ips_well_sums_add");
  if (rc != IPS_OK) return rc;
  if (N < 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_sums_add: negative row count");
  if (N == 0) return IPS_OK;
  if (!rows || !well) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_well_sums_add: NULL rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* sums = reinterpret_cast<double*>(ws);
  int* counts = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + wm_sums_bytes(n_wells, D));
  const int cols = D + 1, subs = WM_THREADS / cols;
  const long long rows_per_block = (long long)subs * WM_ROWS;
  const long long blocks = (N + rows_per_block - 1) / rows_per_block;
  if (blocks > 0x7fffffffLL) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_well_sums_add: too many rows for one call");
  well_accumulate_kernel<<<(unsigned)blocks, WM_THREADS, 0, st>>>(rows, well, sums, counts, N, D, n_wells, cols, subs);
  IPS_LAUNCH_OK("well_accumulate_kernel");
  return IPS_OK;
}

extern "C" int ips_well_sums_finalize(const void* ws, size_t ws_bytes, double* mean_out, int32_t* count_out, int D,
                                      int n_wells, ips_stream_t stream) {
  const int rc = wm_check(ws, ws_bytes, D, n_wells, "ips_well_sums_finalize");
  if (rc != IPS_OK) return rc;
  if (!mean_out || !count_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_well_sums_finalize: NULL output");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const double* sums = reinterpret_cast<const double*>(ws);
  const int* counts = reinterpret_cast<const int*>(reinterpret_cast<const char*>(ws) + wm_sums_bytes(n_wells, D));
  const size_t n_sums = (size_t)n_wells * D;
  well_finalize_kernel<<<(unsigned)((n_sums + 255) / 256), 256, 0, st>>>(sums, counts, mean_out, count_out, n_wells, D);
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
