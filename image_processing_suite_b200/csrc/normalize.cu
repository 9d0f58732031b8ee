// Robust-z normalisation of well profiles against the control wells, and the double sigmoid.
//
// Replaces  pycytominer.normalize(method="mad_robustize", samples=<DMSO wells of the timepoint>)
//           Normalize_CP_ami.py:137-142, Pycyto_pertime.py:84-89   (pycytominer RobustMAD:
//           (x - median_ctrl) / (1.4826 * MAD_ctrl + 1e-18), NaN ignored in the statistics)
//      and  double_sigmoid(x, k=3, alpha=2.3538).abs()
//           Feature_select_cosine_ami.py:22-27, :117-118; Pycyto_pertime.py:13-16
// The profile table is tiny after well aggregation (hundreds of wells x hundreds of features):
// one block per feature column, the control values are sorted in shared memory (bitonic),
// median and MAD come from the sorted arrays, then every well of the column is normalised.
#include <math.h>

#include "ips_common.cuh"

namespace ips {

constexpr int NZ_MAX_CTRL = 2048;   // control wells per column held in shared memory
constexpr double NZ_INF = 1.0e308;

__device__ void bitonic_sort_shared(double* a, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const double x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__device__ double median_sorted(const double* a, int n) {
  if (n == 0) return __longlong_as_double(0x7ff8000000000000ll);
  return (n & 1) ? a[n / 2] : 0.5 * (a[n / 2 - 1] + a[n / 2]);
}

__global__ void __launch_bounds__(256)
mad_robustize_kernel(const double* __restrict__ prof, const uint8_t* __restrict__ is_ctrl,
                     double* __restrict__ out, int W, int D, double mad_scale, double eps) {
  __shared__ double vals[NZ_MAX_CTRL];
  __shared__ int n_s;
  __shared__ double med_s, mad_s;
  const int d = blockIdx.x;
  if (threadIdx.x == 0) n_s = 0;
  __syncthreads();
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    if (is_ctrl[w]) {
      const double v = prof[(size_t)w * D + d];
      if (v == v) {                                     // nanmedian: NaN wells are ignored
        const int k = atomicAdd(&n_s, 1);
        if (k < NZ_MAX_CTRL) vals[k] = v;
      }
    }
  }
  __syncthreads();
  const int n = min(n_s, NZ_MAX_CTRL);
  int p2 = 1;
  while (p2 < n) p2 <<= 1;
  for (int i = n + threadIdx.x; i < p2; i += blockDim.x) vals[i] = NZ_INF;
  __syncthreads();
  bitonic_sort_shared(vals, p2);
  if (threadIdx.x == 0) med_s = median_sorted(vals, n);
  __syncthreads();
  const double med = med_s;
  for (int i = threadIdx.x; i < n; i += blockDim.x) vals[i] = fabs(vals[i] - med);
  __syncthreads();
  bitonic_sort_shared(vals, p2);
  if (threadIdx.x == 0) mad_s = median_sorted(vals, n) * mad_scale;
  __syncthreads();
  const double mad = mad_s;
  for (int w = threadIdx.x; w < W; w += blockDim.x)
    out[(size_t)w * D + d] = (prof[(size_t)w * D + d] - med) / (mad + eps);
}

__global__ void double_sigmoid_abs_kernel(const double* __restrict__ x, double* __restrict__ y, size_t n, int k,
                                          double alpha) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = x[i] / alpha;
  const double rk = pow(r, (double)k);
  const double r2k = pow(r, (double)(2 * k));
  y[i] = fabs(rk / sqrt(1.0 + r2k));
}

}  // namespace ips

using namespace ips;

extern "C" int ips_mad_robustize(const double* profiles, const uint8_t* is_control, double* out, int W, int D,
                                 ips_stream_t stream) {
  if (!profiles || !is_control || !out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_mad_robustize: NULL pointer argument");
  if (W <= 0 || D <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_mad_robustize: bad shape W=%d D=%d", W, D);
  if (W > NZ_MAX_CTRL) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_mad_robustize: at most %d wells per call (got %d)", NZ_MAX_CTRL, W);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  mad_robustize_kernel<<<D, 256, 0, st>>>(profiles, is_control, out, W, D, 1.4826, 1e-18);
  IPS_LAUNCH_OK("mad_robustize_kernel");
  return IPS_OK;
}

extern "C" int ips_double_sigmoid_abs(const double* x, double* y, int64_t n, int k, double alpha, ips_stream_t stream) {
  if (!x || !y) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_double_sigmoid_abs: NULL pointer argument");
  if (n < 0 || k <= 0 || !(alpha != 0.0)) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_double_sigmoid_abs: bad arguments");
  if (n == 0) return IPS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double_sigmoid_abs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, (size_t)n, k, alpha);
  IPS_LAUNCH_OK("double_sigmoid_abs_kernel");
  return IPS_OK;
}
