// K6 -- ring sums of the radial power spectrum (Illumination_QC_mult.py:39-43, :61-68).
//
// The reference folds FFT bin (i, j) to the squared radius min(i, H-1-i)^2 + min(j, W-1-j)^2
// (flip-based fold, :39-43), labels it floor(sqrt(r2)) + 1 (:61) and sums |z| and |z|^2 over
// the labels 2 .. floor(min(H, W) / 8) - 1 with scipy.ndimage.sum (:62-68) -- a label-keyed
// segmented sum, the same shape of reduction as K3.  Here each block owns a band of rows,
// keeps one float64 accumulator pair per ring in shared memory and flushes the band's partial
// sums with float64 global atomics.
#include <math.h>

#include "ips_common.cuh"

namespace ips {

constexpr int RING_THREADS = 256;
constexpr int RING_ROWS = 8;

__global__ void ring_zero_kernel(double* a, double* b, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] = 0.0; b[i] = 0.0; }
}

__global__ void __launch_bounds__(RING_THREADS)
ring_sums_kernel(const double2* __restrict__ spec, double* __restrict__ mag_out,
                 double* __restrict__ pow_out, int n_rings, int H, int W) {
  extern __shared__ double sh[];   // [2][n_rings]
  double* s_mag = sh;
  double* s_pow = sh + n_rings;
  for (int i = threadIdx.x; i < 2 * n_rings; i += RING_THREADS) sh[i] = 0.0;
  __syncthreads();
  const int f = blockIdx.y;
  const int y0 = blockIdx.x * RING_ROWS;
  const int y1 = min(H, y0 + RING_ROWS);
  for (int y = y0; y < y1; ++y) {
    const int di = min(y, H - 1 - y);
    // rows whose folded distance alone exceeds the last ring contribute nothing
    if (di > n_rings + 1) continue;
    const double2* row = spec + ((size_t)f * H + y) * W;
    for (int x = threadIdx.x; x < W; x += RING_THREADS) {
      const int dj = min(x, W - 1 - x);
      const long long r2 = (long long)di * di + (long long)dj * dj;
      int r = (int)sqrt((double)r2);
      while ((long long)r * r > r2) --r;
      while ((long long)(r + 1) * (r + 1) <= r2) ++r;
      const int ring = r + 1 - 2;   // label r + 1, first summed label is 2
      if (ring >= 0 && ring < n_rings) {
        const double2 z = row[x];
        const double m = hypot(z.x, z.y);
        atomicAdd(&s_mag[ring], m);
        atomicAdd(&s_pow[ring], m * m);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_rings; i += RING_THREADS) {
    if (s_mag[i] != 0.0) atomicAdd(&mag_out[(size_t)f * n_rings + i], s_mag[i]);
    if (s_pow[i] != 0.0) atomicAdd(&pow_out[(size_t)f * n_rings + i], s_pow[i]);
  }
}

}  // namespace ips

using namespace ips;

extern "C" int ips_ring_sums(const double* spec_interleaved, double* mag_out, double* pow_out, int n_rings,
                             int F, int H, int W, ips_stream_t stream) {
  if (!spec_interleaved || !mag_out || !pow_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_ring_sums: NULL pointer argument");
  if (F < 0 || H <= 0 || W <= 0 || n_rings <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_ring_sums: bad shape");
  if (n_rings > 2048) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_ring_sums: n_rings must be <= 2048");
  if (!aligned16(spec_interleaved)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_ring_sums: spectrum not 16-byte aligned");
  if (F == 0) return IPS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n = (size_t)F * n_rings;
  ring_zero_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mag_out, pow_out, n);
  IPS_LAUNCH_OK("ring_zero_kernel");
  const size_t smem = (size_t)2 * n_rings * sizeof(double);
  ring_sums_kernel<<<dim3((H + RING_ROWS - 1) / RING_ROWS, F), RING_THREADS, smem, st>>>(
      reinterpret_cast<const double2*>(spec_interleaved), mag_out, pow_out, n_rings, H, W);
  IPS_LAUNCH_OK("ring_sums_kernel");
  return IPS_OK;
}
