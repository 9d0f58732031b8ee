// K5 -- Pillow-exact LANCZOS resize of 16-bit planes ("re-binning" of Image_re-binning.py:18).
//
// Pillow's I;16 resampler is separable: a horizontal pass into a uint16 intermediate
// (rounded), then a vertical pass.  Per output index the window and the normalised double
// weights are those of Pillow's precompute_coeffs; pixels accumulate in double, tap by tap,
// with separate multiply and add (no FMA contraction: the rounding of every partial sum is
// part of the contract), the sum is rounded half away from zero and stored byte-wise: a
// negative value becomes 0, a value above 65535 becomes 0xFF00 | (v & 0xFF) -- Pillow clips
// the two bytes independently (oracle/lanczos.py, verified bit-for-bit against Pillow 12.2.0).
//
// The weights are computed on the HOST with the C library's sin(), exactly as Pillow does
// (the device sin() is not bit-identical to glibc's), cached per (in, out) size and uploaded
// into the caller's workspace on the call's stream.
#include <math.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "ips_common.cuh"

namespace ips {

struct LanczosCoeffs {
  int ksize = 0;
  std::vector<int> bounds;    // [out][2]: first input index, tap count
  std::vector<double> kk;     // [out][ksize]
};

static double sinc_pi(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}

static double lanczos3(double x) {
  if (-3.0 <= x && x < 3.0) return sinc_pi(x) * sinc_pi(x / 3.0);
  return 0.0;
}

// Pillow src/libImaging/Resample.c precompute_coeffs, restated (see oracle/lanczos.py).
static LanczosCoeffs make_coeffs(int in_size, int out_size) {
  LanczosCoeffs c;
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 3.0 * filterscale;
  c.ksize = (int)ceil(support) * 2 + 1;
  c.bounds.assign((size_t)out_size * 2, 0);
  c.kk.assign((size_t)out_size * c.ksize, 0.0);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    const int n = xmax - xmin;
    double* k = &c.kk[(size_t)xx * c.ksize];
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
      const double w = lanczos3((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    if (ww != 0.0)
      for (int x = 0; x < n; ++x) k[x] /= ww;
    c.bounds[(size_t)xx * 2] = xmin;
    c.bounds[(size_t)xx * 2 + 1] = n;
  }
  return c;
}

static const LanczosCoeffs& cached_coeffs(int in_size, int out_size) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, LanczosCoeffs> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(in_size, out_size);
  auto it = cache.find(key);
  if (it == cache.end()) it = cache.emplace(key, make_coeffs(in_size, out_size)).first;
  return it->second;
}

__device__ __forceinline__ uint16_t pil_store_u16(double ss) {
  const long long v = __double2ll_rz(ss >= 0.0 ? ss + 0.5 : ss - 0.5);
  if (v < 0) return 0;
  if (v > 65535) return (uint16_t)(0xFF00u | (unsigned)(v & 0xFF));
  return (uint16_t)v;
}

// along rows: in [P][H][W] -> out [P][H][OW]
__global__ void __launch_bounds__(256)
lanczos_h_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                 const int* __restrict__ bounds, const double* __restrict__ kk, int ksize, int H,
                 int W, int OW) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, p = blockIdx.z;
  if (xx >= OW) return;
  const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
  const double* k = kk + (size_t)xx * ksize;
  const uint16_t* row = in + ((size_t)p * H + y) * W + xmin;
  double ss = 0.0;
  for (int t = 0; t < n; ++t) ss = __dadd_rn(ss, __dmul_rn((double)row[t], k[t]));
  out[((size_t)p * H + y) * OW + xx] = pil_store_u16(ss);
}

// along columns: in [P][H][W] -> out [P][OH][W]
__global__ void __launch_bounds__(256)
lanczos_v_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                 const int* __restrict__ bounds, const double* __restrict__ kk, int ksize, int H,
                 int W, int OH) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int yy = blockIdx.y, p = blockIdx.z;
  if (x >= W) return;
  const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
  const double* k = kk + (size_t)yy * ksize;
  const uint16_t* col = in + ((size_t)p * H + ymin) * W + x;
  double ss = 0.0;
  for (int t = 0; t < n; ++t) ss = __dadd_rn(ss, __dmul_rn((double)col[(size_t)t * W], k[t]));
  out[((size_t)p * OH + yy) * W + x] = pil_store_u16(ss);
}

struct LanczosLayout {
  size_t tmp, hb, hk, vb, vk, total;
};

static LanczosLayout lanczos_layout(int C, int H, int W, int OH, int OW) {
  LanczosLayout L;
  const double sw = (double)W / OW, sh = (double)H / OH;
  const int kw = (int)ceil(3.0 * (sw < 1.0 ? 1.0 : sw)) * 2 + 1;
  const int kh = (int)ceil(3.0 * (sh < 1.0 ? 1.0 : sh)) * 2 + 1;
  size_t o = 0;
  L.tmp = o; o += round_up((size_t)C * H * OW * sizeof(uint16_t), 256);
  L.hb = o; o += round_up((size_t)OW * 2 * sizeof(int), 256);
  L.hk = o; o += round_up((size_t)OW * kw * sizeof(double), 256);
  L.vb = o; o += round_up((size_t)OH * 2 * sizeof(int), 256);
  L.vk = o; o += round_up((size_t)OH * kh * sizeof(double), 256);
  L.total = o;
  return L;
}

}  // namespace ips

using namespace ips;

extern "C" size_t ips_lanczos_workspace_bytes(int C, int H, int W, int outH, int outW) {
  if (C <= 0 || H <= 0 || W <= 0 || outH <= 0 || outW <= 0) return 0;
  return lanczos_layout(C, H, W, outH, outW).total;
}

extern "C" int ips_lanczos_resize_u16(const uint16_t* in, uint16_t* out, int C, int H, int W, int outH,
                                      int outW, void* ws, size_t ws_bytes, ips_stream_t stream) {
  if (!in || !out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_lanczos_resize_u16: NULL pointer argument");
  if (C <= 0 || H <= 0 || W <= 0 || outH <= 0 || outW <= 0)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_lanczos_resize_u16: bad shape C=%d %dx%d -> %dx%d", C, H, W, outH, outW);
  if (C > 65535 || H > 65535 || outH > 65535)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_lanczos_resize_u16: C, H and outH must be <= 65535");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool need_h = outW != W, need_v = outH != H;
  if (!need_h && !need_v) {
    IPS_CUDA_OK(cudaMemcpyAsync(out, in, (size_t)C * H * W * sizeof(uint16_t), cudaMemcpyDeviceToDevice, st));
    return IPS_OK;
  }
  const LanczosLayout L = lanczos_layout(C, H, W, outH, outW);
  if (ws == nullptr || ws_bytes < L.total)
    IPS_FAIL(IPS_ERR_NOMEM, "ips_lanczos_resize_u16: needs %zu workspace bytes (got %zu)", L.total, ws_bytes);
  if (!aligned16(ws)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_lanczos_resize_u16: workspace not 16-byte aligned");
  char* base = reinterpret_cast<char*>(ws);
  uint16_t* tmp = reinterpret_cast<uint16_t*>(base + L.tmp);
  const uint16_t* v_in = in;
  int v_W = W;
  if (need_h) {
    const LanczosCoeffs& c = cached_coeffs(W, outW);
    int* db = reinterpret_cast<int*>(base + L.hb);
    double* dk = reinterpret_cast<double*>(base + L.hk);
    IPS_CUDA_OK(cudaMemcpyAsync(db, c.bounds.data(), c.bounds.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    IPS_CUDA_OK(cudaMemcpyAsync(dk, c.kk.data(), c.kk.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    uint16_t* h_out = need_v ? tmp : out;
    lanczos_h_kernel<<<dim3((outW + 255) / 256, H, C), 256, 0, st>>>(in, h_out, db, dk, c.ksize, H, W, outW);
    IPS_LAUNCH_OK("lanczos_h_kernel");
    v_in = h_out;
    v_W = outW;
  }
  if (need_v) {
    const LanczosCoeffs& c = cached_coeffs(H, outH);
    int* db = reinterpret_cast<int*>(base + L.vb);
    double* dk = reinterpret_cast<double*>(base + L.vk);
    IPS_CUDA_OK(cudaMemcpyAsync(db, c.bounds.data(), c.bounds.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    IPS_CUDA_OK(cudaMemcpyAsync(dk, c.kk.data(), c.kk.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    lanczos_v_kernel<<<dim3((v_W + 255) / 256, outH, C), 256, 0, st>>>(v_in, out, db, dk, c.ksize, H, v_W, outH);
    IPS_LAUNCH_OK("lanczos_v_kernel");
  }
  return IPS_OK;
}
