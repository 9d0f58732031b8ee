// K2 -- per-plate illumination-function estimation.
//
// Produces the {ch}_illum.npy functions that Illumination_QC_mult.py:186-193 and
// Cellpose_GPU_s3fs.py:56 load; the reference computes them outside the repository, so the
// definition is the builder's (restated in oracle/illum.py):
//   raw  = mean over the plate's max-projected fields (exact uint32 sums), or per-pixel median
//   sm   = Gaussian(raw, sigma, zero padded, truncate 4.0) / Gaussian(ones)      (separable)
//   m    = sorted(sm[sm > 0])[int(n_pos * robust_frac)];  out = max(sm, m) / m
//
// Kernels:
//   illum_accumulate_kernel  streaming: 8 pixels (one 128-bit word) per thread per field, all
//                            fields of the batch folded in registers before touching acc.
//   illum_median_kernel      per-pixel exact median over a device-resident stack by 16-step
//                            bitwise bisection on the uint16 value (both middle ranks at once).
//   gauss_rows_kernel        one pass of the separable filter along rows: the row segment plus
//                            halo is staged in shared memory, taps accumulate in float64, the
//                            result is divided by the partial weight sum (edge normalisation)
//                            and written TRANSPOSED through a shared tile, so running the same
//                            kernel twice filters both axes and restores the layout.
//   select_* kernels         exact k-th smallest positive float by 4 x 8-bit radix select.
//   rescale_kernel           out = max(sm, m) / m.
#include <math.h>

#include "ips_common.cuh"

namespace ips {

// ---------------------------------------------------------------------------------------
// accumulate
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
illum_accumulate_kernel(const uint16_t* __restrict__ maxproj, uint32_t* __restrict__ acc, int F,
                        size_t n_words /* C*H*W/8 */) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_words) return;
  const uint64_t pol = policy_evict_first();
  uint32_t s[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = 0u;
  const uint4* src = reinterpret_cast<const uint4*>(maxproj) + i;
  int f = 0;
  for (; f + 4 <= F; f += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ldg128_stream(src + (size_t)(f + u) * n_words, pol);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t x[8];
      unpack_u16x8(v[u], x);
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] += x[k];
    }
  }
  for (; f < F; ++f) {
    uint32_t x[8];
    unpack_u16x8(ldg128_stream(src + (size_t)f * n_words, pol), x);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += x[k];
  }
  uint4* a = reinterpret_cast<uint4*>(acc) + 2 * i;
  uint4 a0 = a[0], a1 = a[1];
  a0.x += s[0]; a0.y += s[1]; a0.z += s[2]; a0.w += s[3];
  a1.x += s[4]; a1.y += s[5]; a1.z += s[6]; a1.w += s[7];
  a[0] = a0;
  a[1] = a1;
}

__global__ void illum_accumulate_scalar_kernel(const uint16_t* __restrict__ maxproj,
                                               uint32_t* __restrict__ acc, int F, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = 0;
  for (int f = 0; f < F; ++f) s += maxproj[(size_t)f * n + i];
  acc[i] += s;
}

__global__ void illum_mean_kernel(const uint32_t* __restrict__ acc, float* __restrict__ raw, double inv_n,
                                  size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) raw[i] = (float)((double)acc[i] * inv_n);
}

// ---------------------------------------------------------------------------------------
// median
// ---------------------------------------------------------------------------------------
// values < cand counted for two ranks at once; k-th smallest = largest v with #(x < v) <= k
__global__ void __launch_bounds__(256)
illum_median_kernel(const uint16_t* __restrict__ stack, float* __restrict__ out, int N, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k_hi = N / 2;
  const int k_lo = (N & 1) ? k_hi : k_hi - 1;
  unsigned lo = 0u, hi = 0u;
  for (int bit = 15; bit >= 0; --bit) {
    const unsigned c_lo = lo | (1u << bit), c_hi = hi | (1u << bit);
    int n_lo = 0, n_hi = 0;
    for (int f = 0; f < N; ++f) {
      const unsigned v = stack[(size_t)f * n + i];
      n_lo += v < c_lo;
      n_hi += v < c_hi;
    }
    if (n_lo <= k_lo) lo = c_lo;
    if (n_hi <= k_hi) hi = c_hi;
  }
  out[i] = 0.5f * (float)(lo + hi);   // exact: lo + hi < 2^17
}

// ---------------------------------------------------------------------------------------
// separable Gaussian
// ---------------------------------------------------------------------------------------
// weights w[0..2R] (normalised, double) followed by edge sums S[0..n-1] for a line of length n
__global__ void gauss_weights_kernel(double* __restrict__ w, int R, double sigma) {
  // single block: weights, then normalise
  __shared__ double total;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = -R; k <= R; ++k) t += exp(-0.5 * (double)k * (double)k / (sigma * sigma));
    total = t;
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= 2 * R; k += blockDim.x) {
    const double d = (double)(k - R);
    w[k] = exp(-0.5 * d * d / (sigma * sigma)) / total;
  }
}

__global__ void gauss_edge_sums_kernel(const double* __restrict__ w, double* __restrict__ S, int R, int n) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  double s = 0.0;
  const int k0 = max(-R, -x), k1 = min(R, n - 1 - x);
  for (int k = k0; k <= k1; ++k) s += w[k + R];
  S[x] = s;
}

constexpr int G_TX = 32;   // outputs along the line per block
constexpr int G_TY = 32;   // lines per block

// in [P][n_lines][n] -> out [P][n][n_lines] (transposed), filtered along the last input axis.
__global__ void __launch_bounds__(G_TX * 8)
gauss_rows_kernel(const float* __restrict__ in, float* __restrict__ out, const double* __restrict__ w,
                  const double* __restrict__ S, int R, int n_lines, int n) {
  extern __shared__ float smem[];
  const int span = G_TX + 2 * R;
  float* tile = smem;                            // [G_TY][span]
  float* res = smem + (size_t)G_TY * span;       // [G_TX][G_TY + 1]
  const int p = blockIdx.z;
  const int x0 = blockIdx.x * G_TX, y0 = blockIdx.y * G_TY;
  const float* src = in + (size_t)p * n_lines * n;
  for (int idx = threadIdx.x; idx < G_TY * span; idx += blockDim.x) {
    const int ly = idx / span, lx = idx - ly * span;
    const int y = y0 + ly, x = x0 + lx - R;
    tile[idx] = (y < n_lines && x >= 0 && x < n) ? src[(size_t)y * n + x] : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & (G_TX - 1);
  for (int ly = threadIdx.x / G_TX; ly < G_TY; ly += blockDim.x / G_TX) {
    const float* row = tile + (size_t)ly * span + tx;
    double a0 = 0.0, a1 = 0.0;
    int k = 0;
    for (; k + 1 <= 2 * R; k += 2) {
      a0 = fma((double)row[k], w[k], a0);
      a1 = fma((double)row[k + 1], w[k + 1], a1);
    }
    if (k <= 2 * R) a0 = fma((double)row[k], w[k], a0);
    const int x = x0 + tx;
    res[tx * (G_TY + 1) + ly] = (x < n) ? (float)((a0 + a1) / S[x]) : 0.f;
  }
  __syncthreads();
  float* dst = out + (size_t)p * n_lines * n;
  for (int idx = threadIdx.x; idx < G_TX * G_TY; idx += blockDim.x) {
    const int lx = idx / G_TY, ly = idx - lx * G_TY;
    const int x = x0 + lx, y = y0 + ly;
    if (x < n && y < n_lines) dst[(size_t)x * n_lines + y] = res[lx * (G_TY + 1) + ly];
  }
}

// ---------------------------------------------------------------------------------------
// exact selection of the k-th smallest positive value (per plane), 4 x 8-bit radix passes
// ---------------------------------------------------------------------------------------
struct SelectState {
  unsigned prefix;      // bits fixed so far (high side)
  unsigned long long k; // rank still to find inside the prefix class
  unsigned long long n_pos;
  float result;
  int pad;
};

__global__ void select_init_kernel(SelectState* st, unsigned* hist, int P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) { st[i].prefix = 0u; st[i].k = 0ull; st[i].n_pos = 0ull; st[i].result = 0.f; }
  if (i < P * 256) hist[i] = 0u;
}

// pass = 0..3 (byte 3 first).  Elements take part if positive and their higher bytes equal prefix.
__global__ void __launch_bounds__(256)
select_hist_kernel(const float* __restrict__ v, const SelectState* __restrict__ st,
                   unsigned* __restrict__ hist, size_t n, int pass) {
  __shared__ unsigned h[256];
  const int p = blockIdx.y;
  h[threadIdx.x] = 0u;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const unsigned prefix = st[p].prefix;
  const unsigned mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
  const float* src = v + (size_t)p * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float x = src[i];
    if (x > 0.f) {
      const unsigned b = __float_as_uint(x);
      if ((b & mask) == prefix) atomicAdd(&h[(b >> shift) & 255u], 1u);
    }
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&hist[p * 256 + threadIdx.x], h[threadIdx.x]);
}

__global__ void select_pick_kernel(SelectState* st, unsigned* hist, int pass, double robust_frac) {
  const int p = blockIdx.x;
  if (threadIdx.x != 0) return;
  unsigned* h = hist + p * 256;
  SelectState s = st[p];
  if (pass == 0) {
    unsigned long long tot = 0;
    for (int b = 0; b < 256; ++b) tot += h[b];
    s.n_pos = tot;
    s.k = (unsigned long long)((double)tot * robust_frac);   // int(len(s) * frac)
    if (tot > 0 && s.k >= tot) s.k = tot - 1;
  }
  const int shift = 24 - 8 * pass;
  if (s.n_pos > 0) {
    unsigned long long run = 0;
    int b = 0;
    for (; b < 256; ++b) {
      if (run + h[b] > s.k) break;
      run += h[b];
    }
    if (b > 255) b = 255;
    s.k -= run;
    s.prefix |= (unsigned)b << shift;
    if (pass == 3) s.result = __uint_as_float(s.prefix);
  }
  st[p] = s;
  for (int b = 0; b < 256; ++b) h[b] = 0u;
}

__global__ void rescale_kernel(const float* __restrict__ sm, const SelectState* __restrict__ st,
                               float* __restrict__ out, size_t n) {
  const int p = blockIdx.y;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SelectState s = st[p];
  const size_t o = (size_t)p * n + i;
  if (s.n_pos == 0) { out[o] = 1.f; return; }
  const float m = s.result;
  out[o] = fmaxf(sm[o], m) / m;
}

static size_t smooth_ws_bytes(int C, int H, int W, int Rmax) {
  const size_t plane = (size_t)H * W;
  size_t b = 0;
  b += round_up((size_t)C * plane * sizeof(float), 256);       // transposed intermediate
  b += round_up((size_t)C * plane * sizeof(float), 256);       // smoothed
  b += round_up((size_t)(2 * Rmax + 1) * sizeof(double), 256); // weights
  b += round_up((size_t)W * sizeof(double), 256);              // edge sums along x
  b += round_up((size_t)H * sizeof(double), 256);              // edge sums along y
  b += round_up((size_t)C * sizeof(SelectState), 256);
  b += round_up((size_t)C * 256 * sizeof(unsigned), 256);
  return b;
}

static int gauss_radius(double sigma) { return (int)(4.0 * sigma + 0.5); }

}  // namespace ips

using namespace ips;

extern "C" int ips_illum_accumulate(const uint16_t* maxproj, uint32_t* acc, int F, int C, int H, int W,
                                    ips_stream_t stream) {
  if (!maxproj || !acc) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_illum_accumulate: NULL pointer argument");
  if (F < 0 || C <= 0 || H <= 0 || W <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_illum_accumulate: bad shape");
  if (F == 0) return IPS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n = (size_t)C * H * W;
  if (n % 8 == 0 && aligned16(maxproj) && aligned16(acc)) {
    const size_t words = n / 8;
    illum_accumulate_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(maxproj, acc, F, words);
  } else {
    illum_accumulate_scalar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(maxproj, acc, F, n);
  }
  IPS_LAUNCH_OK("illum_accumulate_kernel");
  return IPS_OK;
}

extern "C" size_t ips_illum_finalize_workspace_bytes(int C, int H, int W) {
  if (C <= 0 || H <= 0 || W <= 0) return 0;
  // radius is capped by the image: taps beyond max(H, W) never meet a pixel
  const int Rmax = (H > W ? H : W);
  return round_up((size_t)C * H * W * sizeof(float), 256) + smooth_ws_bytes(C, H, W, Rmax);
}

extern "C" int ips_illum_smooth_rescale(const float* raw, double sigma, double robust_frac, float* illum_out,
                                        void* ws, size_t ws_bytes, int C, int H, int W, ips_stream_t stream) {
  if (!raw || !illum_out || !ws) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_illum_smooth_rescale: NULL pointer argument");
  if (C <= 0 || H <= 0 || W <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_illum_smooth_rescale: bad shape");
  if (!(sigma > 0.0)) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_illum_smooth_rescale: sigma must be positive");
  if (!(robust_frac >= 0.0 && robust_frac < 1.0))
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_illum_smooth_rescale: robust_frac must be in [0, 1)");
  const int Rmax = (H > W ? H : W);
  const size_t need = smooth_ws_bytes(C, H, W, Rmax);
  if (ws_bytes < need) IPS_FAIL(IPS_ERR_NOMEM, "ips_illum_smooth_rescale: needs %zu workspace bytes (got %zu)", need, ws_bytes);
  if (!aligned16(ws)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_illum_smooth_rescale: workspace not 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t plane = (size_t)H * W;
  char* q = reinterpret_cast<char*>(ws);
  float* tmp = reinterpret_cast<float*>(q); q += round_up((size_t)C * plane * sizeof(float), 256);
  float* sm = reinterpret_cast<float*>(q); q += round_up((size_t)C * plane * sizeof(float), 256);
  double* w = reinterpret_cast<double*>(q); q += round_up((size_t)(2 * Rmax + 1) * sizeof(double), 256);
  double* Sx = reinterpret_cast<double*>(q); q += round_up((size_t)W * sizeof(double), 256);
  double* Sy = reinterpret_cast<double*>(q); q += round_up((size_t)H * sizeof(double), 256);
  SelectState* sel = reinterpret_cast<SelectState*>(q); q += round_up((size_t)C * sizeof(SelectState), 256);
  unsigned* hist = reinterpret_cast<unsigned*>(q);

  // scipy's radius; taps that can never reach a pixel are dropped (they multiply zeros, and
  // the kernel is normalised over the full radius first, as scipy does)
  const int R_full = gauss_radius(sigma);
  if (R_full > 2 * Rmax) {
    // the normalisation constant still needs the full radius: handled in gauss_weights_kernel
  }
  const int R = R_full;
  if (R > Rmax) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_illum_smooth_rescale: sigma %.3f too large for a %dx%d image (radius %d > %d)", sigma, H, W, R, Rmax);
  gauss_weights_kernel<<<1, 256, 0, st>>>(w, R, sigma);
  IPS_LAUNCH_OK("gauss_weights_kernel");
  gauss_edge_sums_kernel<<<(W + 255) / 256, 256, 0, st>>>(w, Sx, R, W);
  IPS_LAUNCH_OK("gauss_edge_sums_kernel");
  gauss_edge_sums_kernel<<<(H + 255) / 256, 256, 0, st>>>(w, Sy, R, H);
  IPS_LAUNCH_OK("gauss_edge_sums_kernel");
  const size_t smem = ((size_t)G_TY * (G_TX + 2 * R) + (size_t)G_TX * (G_TY + 1)) * sizeof(float);
  if (smem > 220 * 1024) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_illum_smooth_rescale: sigma %.3f needs %zu bytes of shared memory", sigma, smem);
  IPS_CUDA_OK(cudaFuncSetAttribute(gauss_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    dim3 grid((W + G_TX - 1) / G_TX, (H + G_TY - 1) / G_TY, C);
    gauss_rows_kernel<<<grid, G_TX * 8, smem, st>>>(raw, tmp, w, Sx, R, H, W);       // [C][H][W] -> [C][W][H]
    IPS_LAUNCH_OK("gauss_rows_kernel");
    dim3 grid2((H + G_TX - 1) / G_TX, (W + G_TY - 1) / G_TY, C);
    gauss_rows_kernel<<<grid2, G_TX * 8, smem, st>>>(tmp, sm, w, Sy, R, W, H);       // [C][W][H] -> [C][H][W]
    IPS_LAUNCH_OK("gauss_rows_kernel");
  }
  select_init_kernel<<<(C * 256 + 255) / 256, 256, 0, st>>>(sel, hist, C);
  IPS_LAUNCH_OK("select_init_kernel");
  for (int pass = 0; pass < 4; ++pass) {
    const unsigned bx = (unsigned)((plane + 256 * 16 - 1) / (256 * 16));
    select_hist_kernel<<<dim3(bx, C), 256, 0, st>>>(sm, sel, hist, plane, pass);
    IPS_LAUNCH_OK("select_hist_kernel");
    select_pick_kernel<<<C, 32, 0, st>>>(sel, hist, pass, robust_frac);
    IPS_LAUNCH_OK("select_pick_kernel");
  }
  rescale_kernel<<<dim3((unsigned)((plane + 255) / 256), C), 256, 0, st>>>(sm, sel, illum_out, plane);
  IPS_LAUNCH_OK("rescale_kernel");
  return IPS_OK;
}

extern "C" int ips_illum_finalize(const uint32_t* acc, uint64_t n_fields, double sigma, double robust_frac,
                                  float* illum_out, void* ws, size_t ws_bytes, int C, int H, int W,
                                  ips_stream_t stream) {
  if (!acc || !illum_out || !ws) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_illum_finalize: NULL pointer argument");
  if (C <= 0 || H <= 0 || W <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_illum_finalize: bad shape");
  if (n_fields == 0) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_illum_finalize: n_fields is 0");
  const size_t need = ips_illum_finalize_workspace_bytes(C, H, W);
  if (ws_bytes < need) IPS_FAIL(IPS_ERR_NOMEM, "ips_illum_finalize: needs %zu workspace bytes (got %zu)", need, ws_bytes);
  if (!aligned16(ws)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_illum_finalize: workspace not 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n = (size_t)C * H * W;
  float* raw = reinterpret_cast<float*>(ws);
  illum_mean_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(acc, raw, 1.0 / (double)n_fields, n);
  IPS_LAUNCH_OK("illum_mean_kernel");
  const size_t off = round_up(n * sizeof(float), 256);
  return ips_illum_smooth_rescale(raw, sigma, robust_frac, illum_out, reinterpret_cast<char*>(ws) + off,
                                  ws_bytes - off, C, H, W, stream);
}

extern "C" int ips_illum_median(const uint16_t* stack, float* raw_out, int N, int C, int H, int W,
                                ips_stream_t stream) {
  if (!stack || !raw_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_illum_median: NULL pointer argument");
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_illum_median: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n = (size_t)C * H * W;
  illum_median_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(stack, raw_out, N, n);
  IPS_LAUNCH_OK("illum_median_kernel");
  return IPS_OK;
}
