// K6 -- the PowerLogLogSlope metric of Illumination_QC_mult.py without host round trips.
//
// rps() (:31-70) + the regression of calculate_qc_metrics (:104-116), one image per call:
//   1. ips_rps_prepare    float64 image (optionally uint16 / float64 illumination function: the A2
//                         divide of :145-150 in float64) -> mean, min, max (two-level deterministic
//                         reduction), median of |x - mean| by an EXACT radix select on the float64
//                         bit patterns (np.median: mean of the two middle values), then the FFT
//                         input  x / med - mean / med  (or x - mean when the image is constant, :52);
//   2. the FFT itself is the library's (cuFFT through torch.fft.rfft2, as scipy.fftpack is the
//      reference's): real input, so only the Hermitian half [H][W/2+1] is computed;
//   3. ips_ring_sums_half ring-keyed sums of |z| and |z|^2 over that half: every bin is added to its
//                         own ring and, when its mirror bin (H - i, W - j) lies outside the half, to
//                         the mirror's ring -- the two differ because the reference folds radii with
//                         flips, min(j, W - 1 - j), not min(j, W - j) (:39-43);
//   4. ips_loglog_slope   least-squares slope of log(power) over log(ring label) on the rings with
//                         power > 0, 0.0 when fewer than three qualify (:108-114).
#include <math.h>

#include <algorithm>

#include "ips_common.cuh"

namespace ips {

constexpr int QC_THREADS = 256;
constexpr int QC_MAX_BLOCKS = 1184;            // 8 per SM on 148 SMs

struct QcState {
  double mean, vmin, vmax, median;
  unsigned long long prefix;                    // key bits fixed so far
  unsigned long long rank1, rank2;              // remaining ranks inside the current prefix class
  unsigned long long v1_key, next_key;          // lower middle value, smallest key above it
  int need_next, constant;
};

__device__ __forceinline__ double qc_load(const uint16_t* raw, const double* img, const double* illum, size_t i) {
  double x = raw != nullptr ? (double)raw[i] : img[i];
  if (illum != nullptr) x = x / illum[i];
  return x;
}

// pass 1: corrected image (when it is computed here), per-block sum / min / max
__global__ void __launch_bounds__(QC_THREADS)
qc_stats_kernel(const uint16_t* __restrict__ raw, const double* __restrict__ img, const double* __restrict__ illum,
                double* __restrict__ x_out, size_t n, double* __restrict__ part) {
  __shared__ double s_sum[QC_THREADS / 32], s_min[QC_THREADS / 32], s_max[QC_THREADS / 32];
  double sum = 0.0, lo = INFINITY, hi = -INFINITY;
  for (size_t i = (size_t)blockIdx.x * QC_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * QC_THREADS) {
    const double x = qc_load(raw, img, illum, i);
    if (x_out != nullptr) x_out[i] = x;
    sum += x;
    lo = fmin(lo, x);
    hi = fmax(hi, x);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_down_sync(0xffffffffu, sum, o);
    lo = fmin(lo, __shfl_down_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_down_sync(0xffffffffu, hi, o));
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_sum[w] = sum; s_min[w] = lo; s_max[w] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < QC_THREADS / 32; ++k) { sum += s_sum[k]; lo = fmin(lo, s_min[k]); hi = fmax(hi, s_max[k]); }
    part[3 * blockIdx.x] = sum; part[3 * blockIdx.x + 1] = lo; part[3 * blockIdx.x + 2] = hi;
  }
}

__global__ void qc_finish_stats_kernel(const double* __restrict__ part, int blocks, size_t n, QcState* st,
                                       unsigned* __restrict__ hist) {
  if (threadIdx.x == 0) {
    double sum = 0.0, lo = INFINITY, hi = -INFINITY;
    for (int b = 0; b < blocks; ++b) { sum += part[3 * b]; lo = fmin(lo, part[3 * b + 1]); hi = fmax(hi, part[3 * b + 2]); }
    st->mean = sum / (double)n;
    st->vmin = lo; st->vmax = hi;
    st->constant = !(hi - lo > 0.0);            // np.ptp(img) > 0 (:52); NaN images count as constant here
    st->prefix = 0ull;
    st->rank1 = (unsigned long long)((n - 1) / 2);
    st->rank2 = (unsigned long long)(n / 2);
    st->median = 1.0;
    st->need_next = 0;
    st->next_key = ~0ull;
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
}

// one radix pass: histogram of the next 8 key bits over the elements that share the prefix
__global__ void __launch_bounds__(QC_THREADS)
qc_hist_kernel(const double* __restrict__ x, size_t n, const QcState* __restrict__ st, unsigned* __restrict__ hist,
               int pass) {
  __shared__ unsigned sh[256];
  if (st->constant) return;
  sh[threadIdx.x] = 0u;
  __syncthreads();
  const double mean = st->mean;
  const unsigned long long prefix = st->prefix;
  const int shift = 56 - 8 * pass;
  const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
  for (size_t i0 = (size_t)blockIdx.x * QC_THREADS; i0 < n; i0 += (size_t)gridDim.x * QC_THREADS) {
    const size_t i = i0 + threadIdx.x;
    bool in = false;
    unsigned dgt = 0u;
    if (i < n) {
      const unsigned long long key = (unsigned long long)__double_as_longlong(fabs(x[i] - mean));
      in = (key & himask) == prefix;
      dgt = (unsigned)(key >> shift) & 255u;
    }
    // one shared-memory atomic per distinct digit of the warp (the leading bytes are all but constant)
    const unsigned active = __ballot_sync(0xffffffffu, in);
    if (in) {
      const unsigned peers = __match_any_sync(active, dgt);
      if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh[dgt], (unsigned)__popc(peers));
    }
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void qc_pick_kernel(QcState* st, unsigned* hist, int pass) {
  if (threadIdx.x != 0) return;
  if (!st->constant) {
    const int shift = 56 - 8 * pass;
    unsigned long long r1 = st->rank1, acc = 0ull;
    int d = 0;
    for (; d < 256; ++d) {
      if (acc + hist[d] > r1) break;
      acc += hist[d];
    }
    if (d == 256) d = 255;
    // both middle ranks follow the lower one; when the upper falls out of this class the smallest
    // key above the lower middle value is found by one more pass at the end
    const unsigned long long in_bin = hist[d];
    st->rank1 = r1 - acc;
    if (!st->need_next) {
      if (st->rank2 - acc < in_bin) st->rank2 -= acc;
      else st->need_next = 1;
    }
    st->prefix |= (unsigned long long)d << shift;
    if (pass == 7) st->v1_key = st->prefix;
  }
  for (int i = 0; i < 256; ++i) hist[i] = 0u;
}

__global__ void __launch_bounds__(QC_THREADS)
qc_next_kernel(const double* __restrict__ x, size_t n, QcState* st) {
  if (st->constant || !st->need_next) return;
  const double mean = st->mean;
  const unsigned long long v1 = st->v1_key;
  unsigned long long best = ~0ull;
  for (size_t i = (size_t)blockIdx.x * QC_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * QC_THREADS) {
    const unsigned long long key = (unsigned long long)__double_as_longlong(fabs(x[i] - mean));
    if (key > v1 && key < best) best = key;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_down_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
  }
  if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(&st->next_key, best);
}

__global__ void qc_median_kernel(QcState* st) {
  if (threadIdx.x != 0 || st->constant) return;
  const double v1 = __longlong_as_double((long long)st->v1_key);
  const double v2 = st->need_next ? __longlong_as_double((long long)st->next_key) : v1;
  st->median = 0.5 * (v1 + v2);                 // rank1 == rank2 for odd n: the value itself
}

__global__ void __launch_bounds__(QC_THREADS)
qc_normalize_kernel(const double* __restrict__ x, double* __restrict__ z, size_t n, const QcState* __restrict__ st) {
  const size_t i = (size_t)blockIdx.x * QC_THREADS + threadIdx.x;
  if (i >= n) return;
  if (st->constant) z[i] = x[i] - st->mean;                       // no normalisation (:52), DC removed (:57)
  else z[i] = x[i] / st->median - st->mean / st->median;          // img / med, then minus its mean (:53, :57)
}

// ---- ring sums over the Hermitian half -----------------------------------------------------------
constexpr int QCR_ROWS = 8;

__device__ __forceinline__ int qc_ring_of(int di, int dj) {
  const long long r2 = (long long)di * di + (long long)dj * dj;
  int r = (int)sqrt((double)r2);
  while ((long long)r * r > r2) --r;
  while ((long long)(r + 1) * (r + 1) <= r2) ++r;
  return r + 1 - 2;                              // label r + 1, first summed label is 2
}

__global__ void __launch_bounds__(QC_THREADS)
ring_sums_half_kernel(const double2* __restrict__ spec, double* __restrict__ mag_out, double* __restrict__ pow_out,
                      int n_rings, int H, int W, int Wh) {
  extern __shared__ double qsh[];               // [2][n_rings]
  double* s_mag = qsh;
  double* s_pow = qsh + n_rings;
  for (int i = threadIdx.x; i < 2 * n_rings; i += QC_THREADS) qsh[i] = 0.0;
  __syncthreads();
  const int f = blockIdx.y;
  const int y0 = blockIdx.x * QCR_ROWS, y1 = min(H, y0 + QCR_ROWS);
  for (int y = y0; y < y1; ++y) {
    const int ym = (H - y) % H;                 // row of the mirror bin
    const int di = min(y, H - 1 - y), dim_ = min(ym, H - 1 - ym);
    if (di > n_rings + 1 && dim_ > n_rings + 1) continue;
    const double2* row = spec + ((size_t)f * H + y) * Wh;
    for (int x = threadIdx.x; x < Wh; x += QC_THREADS) {
      const int ring = qc_ring_of(di, min(x, W - 1 - x));
      const int xm = (W - x) % W;               // column of the mirror bin; inside the half iff xm < Wh
      const int ringm = xm >= Wh ? qc_ring_of(dim_, min(xm, W - 1 - xm)) : -1;
      const bool a = ring >= 0 && ring < n_rings, b = ringm >= 0 && ringm < n_rings;
      if (a || b) {
        const double2 zc = row[x];
        const double m = hypot(zc.x, zc.y);
        if (a) { atomicAdd(&s_mag[ring], m); atomicAdd(&s_pow[ring], m * m); }
        if (b) { atomicAdd(&s_mag[ringm], m); atomicAdd(&s_pow[ringm], m * m); }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_rings; i += QC_THREADS) {
    if (s_mag[i] != 0.0) atomicAdd(&mag_out[(size_t)f * n_rings + i], s_mag[i]);
    if (s_pow[i] != 0.0) atomicAdd(&pow_out[(size_t)f * n_rings + i], s_pow[i]);
  }
}

__global__ void qc_zero2_kernel(double* a, double* b, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] = 0.0; b[i] = 0.0; }
}

// ---- log-log least squares, one block per image ----------------------------------------------------
__global__ void __launch_bounds__(QC_THREADS)
loglog_slope_kernel(const double* __restrict__ powersum, double* __restrict__ slope_out, int n_rings) {
  __shared__ double red[3][QC_THREADS / 32];
  __shared__ double bc[3];
  const double* p = powersum + (size_t)blockIdx.x * n_rings;
  auto block_sum3 = [&](double a, double b, double c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, o);
      b += __shfl_down_sync(0xffffffffu, b, o);
      c += __shfl_down_sync(0xffffffffu, c, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; red[2][threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int k = 0; k < QC_THREADS / 32; ++k) { s0 += red[0][k]; s1 += red[1][k]; s2 += red[2][k]; }
      bc[0] = s0; bc[1] = s1; bc[2] = s2;
    }
    __syncthreads();
  };
  double cnt = 0.0, sx = 0.0, sy = 0.0;
  for (int i = threadIdx.x; i < n_rings; i += QC_THREADS)
    if (p[i] > 0.0) { cnt += 1.0; sx += log((double)(i + 2)); sy += log(p[i]); }
  block_sum3(cnt, sx, sy);
  const double m = bc[0], xbar = bc[1] / fmax(bc[0], 1.0), ybar = bc[2] / fmax(bc[0], 1.0);
  double sxx = 0.0, sxy = 0.0;
  for (int i = threadIdx.x; i < n_rings; i += QC_THREADS)
    if (p[i] > 0.0) {
      const double dx = log((double)(i + 2)) - xbar;
      sxx += dx * dx;
      sxy += dx * (log(p[i]) - ybar);
    }
  block_sum3(sxx, sxy, 0.0);
  if (threadIdx.x == 0) slope_out[blockIdx.x] = m > 2.0 ? bc[1] / bc[0] : 0.0;
}

}  // namespace ips

using namespace ips;

extern "C" size_t ips_rps_prepare_workspace_bytes(int64_t n) {
  (void)n;
  return round_up(sizeof(QcState), 256) + round_up(256 * sizeof(unsigned), 256) + round_up((size_t)3 * QC_MAX_BLOCKS * sizeof(double), 256);
}

extern "C" int ips_rps_prepare(const uint16_t* raw_u16, const double* img_f64, const double* illum_f64, double* corrected_out,
                               double* fft_in_out, int64_t n, void* ws, size_t ws_bytes, ips_stream_t stream) {
  if ((raw_u16 == nullptr) == (img_f64 == nullptr))
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_rps_prepare: exactly one of the uint16 and the float64 image must be given");
  if (!fft_in_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_rps_prepare: NULL output");
  if (n <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_rps_prepare: empty image");
  const bool computed = raw_u16 != nullptr || illum_f64 != nullptr;    // the float64 image is produced here
  if (computed && corrected_out == nullptr)
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_rps_prepare: a uint16 image or a division needs corrected_out (it may alias fft_in_out)");
  const size_t need = ips_rps_prepare_workspace_bytes(n);
  if (ws == nullptr || ws_bytes < need) IPS_FAIL(IPS_ERR_NOMEM, "ips_rps_prepare: needs %zu workspace bytes (got %zu)", need, ws_bytes);
  if (!aligned16(ws)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_rps_prepare: workspace not 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* w = reinterpret_cast<char*>(ws);
  QcState* state = reinterpret_cast<QcState*>(w);
  unsigned* hist = reinterpret_cast<unsigned*>(w + round_up(sizeof(QcState), 256));
  double* part = reinterpret_cast<double*>(w + round_up(sizeof(QcState), 256) + round_up(256 * sizeof(unsigned), 256));
  const size_t N = (size_t)n;
  const int blocks = (int)std::min<size_t>((N + QC_THREADS - 1) / QC_THREADS, (size_t)QC_MAX_BLOCKS);
  qc_stats_kernel<<<blocks, QC_THREADS, 0, st>>>(raw_u16, img_f64, illum_f64, computed ? corrected_out : nullptr, N, part);
  IPS_LAUNCH_OK("qc_stats_kernel");
  qc_finish_stats_kernel<<<1, 256, 0, st>>>(part, blocks, N, state, hist);
  IPS_LAUNCH_OK("qc_finish_stats_kernel");
  const double* x = computed ? corrected_out : img_f64;
  for (int pass = 0; pass < 8; ++pass) {
    qc_hist_kernel<<<blocks, QC_THREADS, 0, st>>>(x, N, state, hist, pass);
    IPS_LAUNCH_OK("qc_hist_kernel");
    qc_pick_kernel<<<1, 32, 0, st>>>(state, hist, pass);
    IPS_LAUNCH_OK("qc_pick_kernel");
  }
  qc_next_kernel<<<blocks, QC_THREADS, 0, st>>>(x, N, state);
  IPS_LAUNCH_OK("qc_next_kernel");
  qc_median_kernel<<<1, 32, 0, st>>>(state);
  IPS_LAUNCH_OK("qc_median_kernel");
  qc_normalize_kernel<<<(unsigned)((N + QC_THREADS - 1) / QC_THREADS), QC_THREADS, 0, st>>>(x, fft_in_out, N, state);
  IPS_LAUNCH_OK("qc_normalize_kernel");
  return IPS_OK;
}

extern "C" int ips_ring_sums_half(const double* half_spec_interleaved, double* mag_out, double* pow_out, int n_rings, int F,
                                  int H, int W, ips_stream_t stream) {
  if (!half_spec_interleaved || !mag_out || !pow_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_ring_sums_half: NULL pointer argument");
  if (F < 0 || H <= 0 || W <= 0 || n_rings <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_ring_sums_half: bad shape");
  if (n_rings > 2048) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_ring_sums_half: n_rings must be <= 2048");
  if (!aligned16(half_spec_interleaved)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_ring_sums_half: spectrum not 16-byte aligned");
  if (F == 0) return IPS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n = (size_t)F * n_rings;
  qc_zero2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mag_out, pow_out, n);
  IPS_LAUNCH_OK("qc_zero2_kernel");
  const size_t smem = (size_t)2 * n_rings * sizeof(double);
  ring_sums_half_kernel<<<dim3((H + QCR_ROWS - 1) / QCR_ROWS, F), QC_THREADS, smem, st>>>(
      reinterpret_cast<const double2*>(half_spec_interleaved), mag_out, pow_out, n_rings, H, W, W / 2 + 1);
  IPS_LAUNCH_OK("ring_sums_half_kernel");
  return IPS_OK;
}

extern "C" int ips_loglog_slope(const double* powersum, double* slope_out, int n_rings, int F, ips_stream_t stream) {
  if (!powersum || !slope_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_loglog_slope: NULL pointer argument");
  if (n_rings <= 0 || F < 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_loglog_slope: bad shape");
  if (F == 0) return IPS_OK;
  loglog_slope_kernel<<<F, QC_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(powersum, slope_out, n_rings);
  IPS_LAUNCH_OK("loglog_slope_kernel");
  return IPS_OK;
}
