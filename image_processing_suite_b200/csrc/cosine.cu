// K4 -- replicate-group cosine similarity, strict upper triangle.
//
// Replaces  cosine_similarity(features) -> triu(k=1) -> mean
//           Feature_select_cosine_ami.py:145-149, Pycyto_pertime.py:132-140
// Rows are L2-normalised (zero rows stay zero, as scikit-learn does), the Gram matrix of the
// normalised rows is formed tile by tile and never stored: every 64 x 64 tile is masked to
// i < j and group[i] == group[j] in registers and reduced to per-group float64 sums.  Groups
// are contiguous row ranges, so a tile whose row and column ranges share no group is skipped
// before it touches memory -- for the reference's workload (thousands of replicate groups of
// a handful of wells) only the block diagonal is ever computed.
//
// This file is the exact-fp32 CUDA-core path (float32 products, float64 accumulation of the
// tile sums): it is what the parity tests pin to 1e-5.  The tensor-core path for the
// single-huge-group case (config 5) lives in cosine_tc.cu.
#include <stdlib.h>

#include "ips_common.cuh"

namespace ips {

// cosine_tc.cu
size_t cosine_tc_workspace_bytes(int N, int D);
int cosine_tc_launch(const float* X, double* sum_out, int N, int D, void* ws, size_t ws_bytes, cudaStream_t st);

// One group of at least this many rows goes to the tensor cores (IPS_COSINE_TC=0 / 1 forces
// the CUDA-core / tensor-core path for testing).
constexpr int CS_TC_MIN_ROWS = 1024;
static int cosine_tc_mode() {
  static const int m = [] {
    const char* s = getenv("IPS_COSINE_TC");
    return (s && *s) ? atoi(s) : -1;
  }();
  return m;
}

constexpr int CS_TILE = 64;     // rows / cols of the Gram tile per block
constexpr int CS_K = 16;        // depth of one shared-memory stage
constexpr int CS_THREADS = 256; // 16 x 16 threads, 4 x 4 outputs each

// one warp per row: x / ||x||  (||x|| == 0 -> row of zeros)
__global__ void __launch_bounds__(256)
cosine_normalise_kernel(const float* __restrict__ X, float* __restrict__ Xn, int N, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* x = X + (size_t)row * D;
  double ss = 0.0;
  for (int d = lane; d < D; d += 32) {
    const double v = (double)x[d];
    ss = fma(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const double nrm = sqrt(ss);
  const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
  float* y = Xn + (size_t)row * D;
  for (int d = lane; d < D; d += 32) y[d] = (float)((double)x[d] * inv);
}

__global__ void cosine_zero_kernel(double* sums, unsigned long long* counts, int n_groups) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_groups) { sums[i] = 0.0; counts[i] = 0ull; }
}

__global__ void cosine_count_kernel(const int32_t* __restrict__ group, unsigned long long* __restrict__ counts,
                                    int N, int n_groups) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int g = group ? group[i] : 0;
  if (g >= 0 && g < n_groups) atomicAdd(&counts[g], 1ull);
}

__global__ void cosine_pairs_kernel(const unsigned long long* __restrict__ counts,
                                    unsigned long long* __restrict__ npairs, int n_groups) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_groups) {
    const unsigned long long n = counts[i];
    npairs[i] = n * (n - (n > 0 ? 1ull : 0ull)) / 2ull;
  }
}

// upper-triangular tile index -> (bi, bj), bi <= bj, enumerated row by row
__device__ __forceinline__ void triu_tile(long long t, int nt, int& bi, int& bj) {
  // row bi starts at offset bi * nt - bi * (bi - 1) / 2
  double b = (2.0 * nt + 1.0 - sqrt((2.0 * nt + 1.0) * (2.0 * nt + 1.0) - 8.0 * (double)t)) * 0.5;
  long long i = (long long)b;
  if (i < 0) i = 0;
  if (i > nt - 1) i = nt - 1;
  while (i > 0 && i * nt - i * (i - 1) / 2 > t) --i;
  while ((i + 1) * nt - (i + 1) * i / 2 <= t) ++i;
  bi = (int)i;
  bj = (int)(t - (i * nt - i * (i - 1) / 2)) + bi;
}

__global__ void __launch_bounds__(CS_THREADS)
cosine_triu_kernel(const float* __restrict__ Xn, const int32_t* __restrict__ group,
                   double* __restrict__ sum_out, int N, int D, int nt, long long n_tiles, int n_groups,
                   const unsigned long long* __restrict__ grp_start, const unsigned long long* __restrict__ grp_count,
                   const unsigned long long* __restrict__ pair_base, double* __restrict__ pairs_out) {
  __shared__ float As[CS_K][CS_TILE + 4];
  __shared__ float Bs[CS_K][CS_TILE + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    int bi, bj;
    triu_tile(t, nt, bi, bj);
    const int r0 = bi * CS_TILE, c0 = bj * CS_TILE;
    const int r1 = min(N, r0 + CS_TILE) - 1, c1 = min(N, c0 + CS_TILE) - 1;
    if (group != nullptr && group[r1] < group[c0]) continue;   // ranges share no group (ids ascend)
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += CS_K) {
      // 64 rows x 16 depth per operand: 1024 elements, 4 per thread
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = threadIdx.x + e * CS_THREADS;
        const int rr = idx >> 4, kk = idx & 15;
        const int k = k0 + kk;
        As[kk][rr] = (r0 + rr < N && k < D) ? Xn[(size_t)(r0 + rr) * D + k] : 0.f;
        Bs[kk][rr] = (c0 + rr < N && k < D) ? Xn[(size_t)(c0 + rr) * D + k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < CS_K; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
    // strict upper triangle within equal groups -> per-row float64 partials
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty * 4 + i;
      if (r > r1) continue;
      const int gr = group ? group[r] : 0;
      double s = 0.0;
      bool any = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + tx * 4 + j;
        if (c <= c1 && c > r && (group == nullptr || group[c] == gr)) {
          s += (double)acc[i][j];
          any = true;
          if (pairs_out != nullptr && gr >= 0 && gr < n_groups) {
            // row-major strict upper triangle of the group, as np.triu_indices(n, k=1) orders it
            const unsigned long long n = grp_count[gr], li = (unsigned long long)r - grp_start[gr],
                                     lj = (unsigned long long)c - grp_start[gr];
            pairs_out[pair_base[gr] + li * n - li * (li + 1) / 2 + (lj - li - 1)] = (double)acc[i][j];
          }
        }
      }
      if (any && gr >= 0 && gr < n_groups) atomicAdd(&sum_out[gr], s);
    }
  }
}

__global__ void cosine_total_pairs_kernel(unsigned long long* pair_base, const unsigned long long* npairs, int n_groups) {
  if (blockIdx.x == 0 && threadIdx.x == 0) pair_base[n_groups] = pair_base[n_groups - 1] + npairs[n_groups - 1];
}

// group starts (first row), sizes and the offset of every group's pair block; one thread
__global__ void cosine_group_layout_kernel(const int32_t* __restrict__ group, int N, int n_groups,
                                           const unsigned long long* __restrict__ counts,
                                           unsigned long long* __restrict__ grp_start,
                                           unsigned long long* __restrict__ pair_base) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  unsigned long long row = 0, pairs = 0;
  for (int g = 0; g < n_groups; ++g) {
    grp_start[g] = row;
    pair_base[g] = pairs;
    const unsigned long long n = counts[g];
    row += n;
    pairs += n * (n > 0 ? n - 1 : 0) / 2;
  }
  (void)group; (void)N;
}

}  // namespace ips

using namespace ips;

extern "C" size_t ips_cosine_workspace_bytes(int N, int D) {
  if (N <= 0 || D <= 0) return 0;
  const size_t exact = round_up((size_t)N * D * sizeof(float), 256) + round_up((size_t)N * sizeof(unsigned long long), 256);
  const size_t tc = cosine_tc_workspace_bytes(N, D) + 1024 + round_up((size_t)N * sizeof(unsigned long long), 256);
  return exact > tc ? exact : tc;
}

extern "C" int ips_cosine_triu(const float* X, const int32_t* group, int n_groups, double* sum_out,
                               uint64_t* npairs_out, int N, int D, void* ws, size_t ws_bytes,
                               ips_stream_t stream) {
  if (!sum_out || !npairs_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_cosine_triu: NULL output");
  if (N < 0 || D <= 0 || n_groups <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_cosine_triu: bad shape N=%d D=%d groups=%d", N, D, n_groups);
  if (group == nullptr && n_groups != 1) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_cosine_triu: group == NULL means one group");
  if (n_groups > N && N > 0) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_cosine_triu: more groups than rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned long long* np = reinterpret_cast<unsigned long long*>(npairs_out);
  cosine_zero_kernel<<<(n_groups + 255) / 256, 256, 0, st>>>(sum_out, np, n_groups);
  IPS_LAUNCH_OK("cosine_zero_kernel");
  if (N == 0) return IPS_OK;
  if (X == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_cosine_triu: X is NULL");
  const size_t need = ips_cosine_workspace_bytes(N, D);
  if (ws == nullptr || ws_bytes < need) IPS_FAIL(IPS_ERR_NOMEM, "ips_cosine_triu: needs %zu workspace bytes (got %zu)", need, ws_bytes);
  if (!aligned16(ws)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_cosine_triu: workspace not 16-byte aligned");
  const int mode = cosine_tc_mode();
  const bool use_tc = group == nullptr && N >= 2 && (mode == 1 || (mode != 0 && N >= CS_TC_MIN_ROWS));
  if (use_tc) {
    // counts live at the start of the workspace, the 1024-aligned bf16 planes behind them
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(ws);
    cosine_zero_kernel<<<1, 32, 0, st>>>(sum_out, cnt, 1);
    IPS_LAUNCH_OK("cosine_zero_kernel");
    cosine_count_kernel<<<(N + 255) / 256, 256, 0, st>>>(nullptr, cnt, N, 1);
    IPS_LAUNCH_OK("cosine_count_kernel");
    cosine_pairs_kernel<<<1, 32, 0, st>>>(cnt, np, 1);
    IPS_LAUNCH_OK("cosine_pairs_kernel");
    const size_t head = round_up((size_t)N * sizeof(unsigned long long), 256);
    char* planes = reinterpret_cast<char*>(ws) + head;
    planes = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(planes) + 1023) & ~(uintptr_t)1023);
    const size_t left = ws_bytes - (size_t)(planes - reinterpret_cast<char*>(ws));
    return cosine_tc_launch(X, sum_out, N, D, planes, left, st);
  }
  float* Xn = reinterpret_cast<float*>(ws);
  unsigned long long* counts =
      reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) + round_up((size_t)N * D * sizeof(float), 256));
  cosine_normalise_kernel<<<(N + 7) / 8, 256, 0, st>>>(X, Xn, N, D);
  IPS_LAUNCH_OK("cosine_normalise_kernel");
  // counts live in the workspace for the pair formula; npairs_out is written from them
  cosine_zero_kernel<<<(n_groups + 255) / 256, 256, 0, st>>>(sum_out, counts, n_groups);
  IPS_LAUNCH_OK("cosine_zero_kernel");
  cosine_count_kernel<<<(N + 255) / 256, 256, 0, st>>>(group, counts, N, n_groups);
  IPS_LAUNCH_OK("cosine_count_kernel");
  cosine_pairs_kernel<<<(n_groups + 255) / 256, 256, 0, st>>>(counts, np, n_groups);
  IPS_LAUNCH_OK("cosine_pairs_kernel");
  const int nt = (N + CS_TILE - 1) / CS_TILE;
  const long long n_tiles = (long long)nt * (nt + 1) / 2;
  const long long want = n_tiles < (long long)sm_count() * 64 ? n_tiles : (long long)sm_count() * 64;
  cosine_triu_kernel<<<(unsigned)want, CS_THREADS, 0, st>>>(Xn, group, sum_out, N, D, nt, n_tiles, n_groups, nullptr,
                                                            nullptr, nullptr, nullptr);
  IPS_LAUNCH_OK("cosine_triu_kernel");
  return IPS_OK;
}

// Same as ips_cosine_triu on the exact fp32 path, and additionally every pair's similarity:
// pairs_out holds, group after group, the row-major strict upper triangle of the group's
// similarity matrix (what Pycyto_pertime.py:150-155 keeps as `cosine_similarities`);
// pair_offsets_out [n_groups + 1] are the block boundaries.  Rows of a group must be contiguous
// and group ids ascending from 0.  pairs_capacity bounds the number of pair values written.
extern "C" size_t ips_cosine_pairs_workspace_bytes(int N, int D) {
  if (N <= 0 || D <= 0) return 0;
  return round_up((size_t)N * D * sizeof(float), 256) + 3 * round_up((size_t)(N + 1) * sizeof(unsigned long long), 256);
}

extern "C" int ips_cosine_triu_pairs(const float* X, const int32_t* group, int n_groups, double* sum_out,
                                     uint64_t* npairs_out, double* pairs_out, uint64_t* pair_offsets_out,
                                     uint64_t pairs_capacity, int N, int D, void* ws, size_t ws_bytes,
                                     ips_stream_t stream) {
  if (!sum_out || !npairs_out || !pairs_out || !pair_offsets_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_cosine_triu_pairs: NULL output");
  if (N <= 0 || D <= 0 || n_groups <= 0 || n_groups > N)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_cosine_triu_pairs: bad shape N=%d D=%d groups=%d", N, D, n_groups);
  if (X == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_cosine_triu_pairs: X is NULL");
  if (group == nullptr && n_groups != 1) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_cosine_triu_pairs: group == NULL means one group");
  // upper bound of the pair count: all rows in one group
  if ((unsigned long long)N * (unsigned long long)(N - 1) / 2 > pairs_capacity && n_groups == 1)
    IPS_FAIL(IPS_ERR_NOMEM, "ips_cosine_triu_pairs: pairs_out too small for %d rows in one group", N);
  const size_t need = ips_cosine_pairs_workspace_bytes(N, D);
  if (ws == nullptr || ws_bytes < need) IPS_FAIL(IPS_ERR_NOMEM, "ips_cosine_triu_pairs: needs %zu workspace bytes (got %zu)", need, ws_bytes);
  if (!aligned16(ws)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_cosine_triu_pairs: workspace not 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* q = reinterpret_cast<char*>(ws);
  float* Xn = reinterpret_cast<float*>(q); q += round_up((size_t)N * D * sizeof(float), 256);
  unsigned long long* counts = reinterpret_cast<unsigned long long*>(q); q += round_up((size_t)(N + 1) * sizeof(unsigned long long), 256);
  unsigned long long* grp_start = reinterpret_cast<unsigned long long*>(q); q += round_up((size_t)(N + 1) * sizeof(unsigned long long), 256);
  unsigned long long* np = reinterpret_cast<unsigned long long*>(npairs_out);
  unsigned long long* pair_base = reinterpret_cast<unsigned long long*>(pair_offsets_out);
  cosine_zero_kernel<<<(n_groups + 255) / 256, 256, 0, st>>>(sum_out, counts, n_groups);
  IPS_LAUNCH_OK("cosine_zero_kernel");
  cosine_count_kernel<<<(N + 255) / 256, 256, 0, st>>>(group, counts, N, n_groups);
  IPS_LAUNCH_OK("cosine_count_kernel");
  cosine_pairs_kernel<<<(n_groups + 255) / 256, 256, 0, st>>>(counts, np, n_groups);
  IPS_LAUNCH_OK("cosine_pairs_kernel");
  cosine_group_layout_kernel<<<1, 32, 0, st>>>(group, N, n_groups, counts, grp_start, pair_base);
  IPS_LAUNCH_OK("cosine_group_layout_kernel");
  // the closing boundary pair_offsets_out[n_groups] = total pairs
  cosine_total_pairs_kernel<<<1, 32, 0, st>>>(pair_base, np, n_groups);
  IPS_LAUNCH_OK("cosine_total_pairs_kernel");
  cosine_normalise_kernel<<<(N + 7) / 8, 256, 0, st>>>(X, Xn, N, D);
  IPS_LAUNCH_OK("cosine_normalise_kernel");
  const int nt = (N + CS_TILE - 1) / CS_TILE;
  const long long n_tiles = (long long)nt * (nt + 1) / 2;
  const long long want = n_tiles < (long long)sm_count() * 64 ? n_tiles : (long long)sm_count() * 64;
  (void)pairs_capacity;
  cosine_triu_kernel<<<(unsigned)want, CS_THREADS, 0, st>>>(Xn, group, sum_out, N, D, nt, n_tiles, n_groups, grp_start,
                                                            counts, pair_base, pairs_out);
  IPS_LAUNCH_OK("cosine_triu_kernel");
  return IPS_OK;
}
