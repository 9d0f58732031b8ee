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
#include <string.h>

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
  // Integer decimation: outputs [u_lo, u_hi) all use the same `u_taps` weights (bit for bit) on
  // windows that advance by `u_step` input pixels: first input of output o is u_base + u_step * o.
  int u_lo = 0, u_hi = 0, u_step = 0, u_taps = 0, u_base = 0;
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
  // longest run of outputs with identical weights and evenly advancing windows (checked, not assumed)
  if (in_size % out_size == 0 && in_size / out_size >= 2) {
    const int step = in_size / out_size;
    int best_lo = 0, best_hi = 0;
    for (int lo = 0; lo < out_size;) {
      int hi = lo + 1;
      const int n0 = c.bounds[(size_t)lo * 2 + 1];
      const double* k0 = &c.kk[(size_t)lo * c.ksize];
      while (hi < out_size && c.bounds[(size_t)hi * 2 + 1] == n0 &&
             c.bounds[(size_t)hi * 2] == c.bounds[(size_t)lo * 2] + step * (hi - lo) &&
             memcmp(&c.kk[(size_t)hi * c.ksize], k0, (size_t)n0 * sizeof(double)) == 0)
        ++hi;
      if (hi - lo > best_hi - best_lo) {
        best_lo = lo;
        best_hi = hi;
      }
      lo = hi;
    }
    if (best_hi - best_lo >= 64) {
      c.u_lo = best_lo;
      c.u_hi = best_hi;
      c.u_step = step;
      c.u_taps = c.bounds[(size_t)best_lo * 2 + 1];
      c.u_base = c.bounds[(size_t)best_lo * 2] - step * best_lo;
    }
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

// |ss| <= 65535 * sum|w| < 2^18: the conversion never leaves int32
__device__ __forceinline__ uint16_t pil_store_u16(double ss) {
  const int v = __double2int_rz(ss >= 0.0 ? ss + 0.5 : ss - 0.5);
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

// ---- shared-memory staged passes ---------------------------------------------------------------
// The two kernels above issue one scalar load and one int -> double conversion per tap.  Here a
// block first converts the input span it needs to double in shared memory (every input pixel
// is converted once), then each thread accumulates its outputs tap by tap from shared memory in
// exactly the same order (same __dmul_rn / __dadd_rn sequence, so the results are identical).
constexpr int LZ_TX = 128;   // outputs along the filtered axis (h pass) / columns (v pass) per block
constexpr int LZ_RH = 8;     // rows per block in the horizontal pass
constexpr int LZ_RV = 8;     // output rows per block in the vertical pass
constexpr int LZ_SPAN_ITERS = 3;   // the horizontal pass stages spans of up to LZ_SPAN_ITERS * LZ_TX pixels

__global__ void __launch_bounds__(LZ_TX)
lanczos_h_staged_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                        const int* __restrict__ bounds, const double* __restrict__ kk, int ksize, int H,
                        int W, int OW, int span_max) {
  extern __shared__ double lz_s[];               // [LZ_RH][span_max] pixels, then [LZ_TX][ksize] weights
  double* kk_s = lz_s + (size_t)LZ_RH * span_max;
  const int p = blockIdx.z;
  const int y0 = blockIdx.y * LZ_RH;
  const int xo0 = blockIdx.x * LZ_TX;
  const int xo1 = min(OW, xo0 + LZ_TX);
  const int lo = bounds[2 * xo0];
  const int hi = bounds[2 * (xo1 - 1)] + bounds[2 * (xo1 - 1) + 1];
  const int span = hi - lo;
  // every load of the block's input span is issued before the first conversion (the loop below
  // is fully unrolled: LZ_RH x LZ_SPAN_ITERS independent loads in flight per thread)
  uint16_t v[LZ_RH][LZ_SPAN_ITERS];
#pragma unroll
  for (int r = 0; r < LZ_RH; ++r) {
    const int y = y0 + r;
    const uint16_t* src = in + ((size_t)p * H + (y < H ? y : H - 1)) * W + lo;
#pragma unroll
    for (int j = 0; j < LZ_SPAN_ITERS; ++j) {
      const int i = threadIdx.x + j * LZ_TX;
      v[r][j] = i < span ? src[i] : (uint16_t)0;
    }
  }
#pragma unroll
  for (int r = 0; r < LZ_RH; ++r)
#pragma unroll
    for (int j = 0; j < LZ_SPAN_ITERS; ++j) {
      const int i = threadIdx.x + j * LZ_TX;
      if (i < span) lz_s[r * span_max + i] = (double)v[r][j];
    }
  // the weights of the block's outputs are one contiguous piece of the table
  const double* kblk = kk + (size_t)xo0 * ksize;
#pragma unroll 8
  for (int idx = threadIdx.x; idx < (xo1 - xo0) * ksize; idx += LZ_TX) kk_s[idx] = kblk[idx];
  __syncthreads();
  const int xx = xo0 + threadIdx.x;
  if (xx >= xo1) return;
  const int xmin = bounds[2 * xx] - lo, n = bounds[2 * xx + 1];
  const double* k = kk_s + (size_t)threadIdx.x * ksize;
  double ss[LZ_RH];
#pragma unroll
  for (int r = 0; r < LZ_RH; ++r) ss[r] = 0.0;
#pragma unroll 4
  for (int t = 0; t < n; ++t) {
    const double kt = k[t];
#pragma unroll
    for (int r = 0; r < LZ_RH; ++r) ss[r] = __dadd_rn(ss[r], __dmul_rn(lz_s[r * span_max + xmin + t], kt));
  }
#pragma unroll
  for (int r = 0; r < LZ_RH; ++r)
    if (y0 + r < H) out[((size_t)p * H + y0 + r) * OW + xx] = pil_store_u16(ss[r]);
}

__global__ void __launch_bounds__(LZ_TX)
lanczos_v_staged_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                        const int* __restrict__ bounds, const double* __restrict__ kk, int ksize, int H,
                        int W, int OH, int span_max) {
  extern __shared__ double lz_s[];               // [span_max][LZ_TX] pixels, then [LZ_RV][ksize] weights
  double* kk_s = lz_s + (size_t)span_max * LZ_TX;
  const int p = blockIdx.z;
  const int yy0 = blockIdx.y * LZ_RV;
  const int yy1 = min(OH, yy0 + LZ_RV);
  const int x = blockIdx.x * LZ_TX + threadIdx.x;
  const int lo = bounds[2 * yy0];
  const int hi = bounds[2 * (yy1 - 1)] + bounds[2 * (yy1 - 1) + 1];
  const int span = hi - lo;
  if (x < W) {
    const uint16_t* col = in + ((size_t)p * H + lo) * W + x;
#pragma unroll 8
    for (int i = 0; i < span; ++i) lz_s[i * LZ_TX + threadIdx.x] = (double)col[(size_t)i * W];
  }
  for (int idx = threadIdx.x; idx < (yy1 - yy0) * ksize; idx += LZ_TX) kk_s[idx] = kk[(size_t)yy0 * ksize + idx];
  __syncthreads();
  if (x >= W) return;
  for (int yy = yy0; yy < yy1; ++yy) {
    const int ymin = bounds[2 * yy] - lo, n = bounds[2 * yy + 1];
    const double* k = kk_s + (size_t)(yy - yy0) * ksize;
    double ss = 0.0;
#pragma unroll 4
    for (int t = 0; t < n; ++t) ss = __dadd_rn(ss, __dmul_rn(lz_s[(ymin + t) * LZ_TX + threadIdx.x], k[t]));
    out[((size_t)p * OH + yy) * W + x] = pil_store_u16(ss);
  }
}


// ---- integer decimation (2160 -> 1080 / 540, the script's resolutions) -------------------------
// When every interior output uses the same TAPS weights on windows STEP pixels apart
// (LanczosCoeffs::u_*, verified on the host), the weights live in registers and each input pixel
// is loaded and converted once: a thread walks its NO * STEP + TAPS - STEP inputs in order and
// adds pixel * weight into every output whose window holds it -- per output the same products
// in the same order as the generic kernels, so the results are bit-identical.  No shared memory;
// HBM sees the input and the output once.
template <int STEP, int TAPS>
struct LzWeights {
  double w[TAPS];
};

// the few outputs outside the uniform interior (clipped windows at the image border, and the
// remainder of the last group): their windows and weights travel as kernel parameters and spare
// threads / block rows of the same launch compute them, so the integer-decimation path needs no
// coefficient upload and no extra launch
constexpr int LZ_EDGE_MAX = 16, LZ_EDGE_TAPS = 25;
struct LzEdges {
  int count;
  int index[LZ_EDGE_MAX], first[LZ_EDGE_MAX], taps[LZ_EDGE_MAX];
  double w[LZ_EDGE_MAX][LZ_EDGE_TAPS];
};

// rows: thread = NO consecutive outputs of one row, lanes side by side along the row.  The block's
// input span is read once with coalesced loads, converted once and parked in shared memory (one
// pad per thread chunk: conflict-free); threads behind the last group take the edge outputs.
template <int STEP, int TAPS, int NO>
__global__ void __launch_bounds__(128)
lanczos_h_uniform_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, const LzWeights<STEP, TAPS> wt,
                         const LzEdges e, int H, int W, int OW, int u_lo, int u_base, int groups) {
  constexpr int CH = NO * STEP;                       // inputs a thread advances by
  constexpr int NIN = CH + TAPS - STEP;               // inputs a thread reads
  constexpr int SPAN = 128 * CH + TAPS - STEP;        // inputs a block reads
  __shared__ double px_s[SPAN + SPAN / CH + 1];
  const int g0 = blockIdx.x * 128;
  const int y = blockIdx.y, p = blockIdx.z;
  const uint16_t* row = in + ((size_t)p * H + y) * W;
  uint16_t* orow = out + ((size_t)p * H + y) * OW;
  const int here = min(128, groups - g0);             // groups of this block (<= 0: edge-only block)
  if (here > 0) {
    const uint16_t* src = row + u_base + STEP * (u_lo + g0 * NO);
    const int span = here * CH + TAPS - STEP;
    for (int i = threadIdx.x; i < span; i += 128) px_s[i + i / CH] = (double)src[i];
  }
  __syncthreads();
  const int g = g0 + threadIdx.x;
  if (g < groups) {
    double ss[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) ss[o] = 0.0;
    const double* mine = px_s + threadIdx.x * (CH + 1);
#pragma unroll
    for (int i = 0; i < NIN; ++i) {
      const double px = mine[i + i / CH];
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        const int t = i - STEP * o;
        if (t >= 0 && t < TAPS) ss[o] = __dadd_rn(ss[o], __dmul_rn(px, wt.w[t]));
      }
    }
    uint16_t* dst = orow + u_lo + g * NO;
#pragma unroll
    for (int o = 0; o < NO; ++o) dst[o] = pil_store_u16(ss[o]);
  } else if (g - groups < e.count) {
    const int k = g - groups;
    const uint16_t* src = row + e.first[k];
    double ss = 0.0;
    for (int t = 0; t < e.taps[k]; ++t) ss = __dadd_rn(ss, __dmul_rn((double)src[t], e.w[k][t]));
    orow[e.index[k]] = pil_store_u16(ss);
  }
}

// columns: thread = NO consecutive output rows of two adjacent columns, lanes side by side along
// the row (coalesced); block rows behind the last group take the edge rows
template <int STEP, int TAPS, int NO, bool PAIRED>
__global__ void __launch_bounds__(128)
lanczos_v_uniform_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, const LzWeights<STEP, TAPS> wt,
                         const LzEdges e, int H, int W, int OH, int u_lo, int u_base, int groups) {
  const int x = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int p = blockIdx.z;
  if (x >= W) return;
  const bool pair = x + 1 < W;
  if ((int)blockIdx.y >= groups) {                     // an edge row
    const int k = blockIdx.y - groups;
    const uint16_t* col = in + ((size_t)p * H + e.first[k]) * W + x;
    double a = 0.0, b = 0.0;
    for (int t = 0; t < e.taps[k]; ++t) {
      a = __dadd_rn(a, __dmul_rn((double)col[(size_t)t * W], e.w[k][t]));
      if (pair) b = __dadd_rn(b, __dmul_rn((double)col[(size_t)t * W + 1], e.w[k][t]));
    }
    uint16_t* dst = out + ((size_t)p * OH + e.index[k]) * W + x;
    dst[0] = pil_store_u16(a);
    if (pair) dst[1] = pil_store_u16(b);
    return;
  }
  const int o0 = u_lo + blockIdx.y * NO;
  const uint16_t* src = in + ((size_t)p * H + u_base + STEP * o0) * W + x;
  constexpr int NIN = NO * STEP + TAPS - STEP;
  double s0[NO], s1[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) s0[o] = s1[o] = 0.0;
#pragma unroll
  for (int i = 0; i < NIN; ++i) {
    const uint16_t* q = src + (size_t)i * W;
    double a, b = 0.0;
    if (PAIRED) {                                      // W even and the planes 4-byte aligned: one load for both columns
      const uint32_t v = *reinterpret_cast<const uint32_t*>(q);
      a = (double)(v & 0xFFFFu);
      b = (double)(v >> 16);
    } else {
      a = (double)q[0];
      if (pair) b = (double)q[1];
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      const int t = i - STEP * o;
      if (t >= 0 && t < TAPS) {
        s0[o] = __dadd_rn(s0[o], __dmul_rn(a, wt.w[t]));
        s1[o] = __dadd_rn(s1[o], __dmul_rn(b, wt.w[t]));
      }
    }
  }
  uint16_t* dst = out + ((size_t)p * OH + o0) * W + x;
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    if (PAIRED) {
      *reinterpret_cast<uint32_t*>(dst + (size_t)o * W) = (uint32_t)pil_store_u16(s0[o]) | ((uint32_t)pil_store_u16(s1[o]) << 16);
    } else {
      dst[(size_t)o * W] = pil_store_u16(s0[o]);
      if (pair) dst[(size_t)o * W + 1] = pil_store_u16(s1[o]);
    }
  }
}

// largest input span any block of `per_block` consecutive outputs needs
static int lanczos_span_max(const LanczosCoeffs& c, int n_out, int per_block) {
  int m = 1;
  for (int o0 = 0; o0 < n_out; o0 += per_block) {
    const int o1 = (o0 + per_block < n_out ? o0 + per_block : n_out) - 1;
    const int span = c.bounds[(size_t)o1 * 2] + c.bounds[(size_t)o1 * 2 + 1] - c.bounds[(size_t)o0 * 2];
    if (span > m) m = span;
  }
  return m;
}

constexpr size_t LZ_SMEM_LIMIT = 96 * 1024;   // keeps at least two blocks per SM

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


template <int STEP, int TAPS>
static LzWeights<STEP, TAPS> uniform_weights(const LanczosCoeffs& c) {
  LzWeights<STEP, TAPS> w;
  for (int t = 0; t < TAPS; ++t) w.w[t] = c.kk[(size_t)c.u_lo * c.ksize + t];
  return w;
}

// outputs [0, lo) and [hi, n_out) as kernel parameters; false when they do not fit
static bool make_edges(const LanczosCoeffs& c, int lo, int hi, int n_out, LzEdges* e) {
  e->count = 0;
  for (int o = 0; o < n_out; ++o) {
    if (o >= lo && o < hi) {
      o = hi - 1;
      continue;
    }
    const int n = c.bounds[(size_t)o * 2 + 1];
    if (e->count == LZ_EDGE_MAX || n > LZ_EDGE_TAPS) return false;
    e->index[e->count] = o;
    e->first[e->count] = c.bounds[(size_t)o * 2];
    e->taps[e->count] = n;
    for (int t = 0; t < n; ++t) e->w[e->count][t] = c.kk[(size_t)o * c.ksize + t];
    e->count++;
  }
  return true;
}

// interior outputs and edge outputs in one launch; false = not applicable
template <int STEP, int TAPS, int NO>
static bool uniform_h_launch(const LanczosCoeffs& c, const uint16_t* in, uint16_t* out, int C, int H, int W, int OW,
                             cudaStream_t st) {
  const int groups = (c.u_hi - c.u_lo) / NO;
  LzEdges e;
  if (groups <= 0 || H > 65535 || !make_edges(c, c.u_lo, c.u_lo + groups * NO, OW, &e)) return false;
  lanczos_h_uniform_kernel<STEP, TAPS, NO><<<dim3((groups + e.count + 127) / 128, H, C), 128, 0, st>>>(
      in, out, uniform_weights<STEP, TAPS>(c), e, H, W, OW, c.u_lo, c.u_base, groups);
  count_launch();
  return true;
}

template <int STEP, int TAPS, int NO>
static bool uniform_v_launch(const LanczosCoeffs& c, const uint16_t* in, uint16_t* out, int C, int H, int W, int OH,
                             cudaStream_t st) {
  const int groups = (c.u_hi - c.u_lo) / NO;
  LzEdges e;
  if (groups <= 0 || groups + LZ_EDGE_MAX > 65535 || !make_edges(c, c.u_lo, c.u_lo + groups * NO, OH, &e)) return false;
  const int pairs = (W + 1) / 2;
  const dim3 grid((pairs + 127) / 128, groups + e.count, C);
  const bool paired = W % 2 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 3u) == 0;
  if (paired)
    lanczos_v_uniform_kernel<STEP, TAPS, NO, true><<<grid, 128, 0, st>>>(in, out, uniform_weights<STEP, TAPS>(c), e, H, W, OH,
                                                                        c.u_lo, c.u_base, groups);
  else
    lanczos_v_uniform_kernel<STEP, TAPS, NO, false><<<grid, 128, 0, st>>>(in, out, uniform_weights<STEP, TAPS>(c), e, H, W, OH,
                                                                         c.u_lo, c.u_base, groups);
  count_launch();
  return true;
}

static bool uniform_h(const LanczosCoeffs& c, const uint16_t* in, uint16_t* out, int C, int H, int W, int OW,
                      cudaStream_t st) {
  if (c.u_step == 2 && c.u_taps == 12) return uniform_h_launch<2, 12, 8>(c, in, out, C, H, W, OW, st);
  if (c.u_step == 3 && c.u_taps == 18) return uniform_h_launch<3, 18, 4>(c, in, out, C, H, W, OW, st);
  if (c.u_step == 4 && c.u_taps == 24) return uniform_h_launch<4, 24, 4>(c, in, out, C, H, W, OW, st);
  return false;
}
static bool uniform_v(const LanczosCoeffs& c, const uint16_t* in, uint16_t* out, int C, int H, int W, int OH,
                      cudaStream_t st) {
  if (c.u_step == 2 && c.u_taps == 12) return uniform_v_launch<2, 12, 8>(c, in, out, C, H, W, OH, st);
  if (c.u_step == 3 && c.u_taps == 18) return uniform_v_launch<3, 18, 4>(c, in, out, C, H, W, OH, st);
  if (c.u_step == 4 && c.u_taps == 24) return uniform_v_launch<4, 24, 4>(c, in, out, C, H, W, OH, st);
  return false;
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
    uint16_t* h_out = need_v ? tmp : out;
    if (!uniform_h(c, in, h_out, C, H, W, outW, st)) {     // integer decimation needs no tables on the device
      int* db = reinterpret_cast<int*>(base + L.hb);
      double* dk = reinterpret_cast<double*>(base + L.hk);
      IPS_CUDA_OK(cudaMemcpyAsync(db, c.bounds.data(), c.bounds.size() * sizeof(int), cudaMemcpyHostToDevice, st));
      IPS_CUDA_OK(cudaMemcpyAsync(dk, c.kk.data(), c.kk.size() * sizeof(double), cudaMemcpyHostToDevice, st));
      const int span_max = lanczos_span_max(c, outW, LZ_TX);
      const size_t smem = ((size_t)LZ_RH * span_max + (size_t)LZ_TX * c.ksize) * sizeof(double);
      if (smem <= LZ_SMEM_LIMIT && span_max <= LZ_SPAN_ITERS * LZ_TX && (H + LZ_RH - 1) / LZ_RH <= 65535) {
        IPS_CUDA_OK(cudaFuncSetAttribute(lanczos_h_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lanczos_h_staged_kernel<<<dim3((outW + LZ_TX - 1) / LZ_TX, (H + LZ_RH - 1) / LZ_RH, C), LZ_TX, smem, st>>>(
            in, h_out, db, dk, c.ksize, H, W, outW, span_max);
        IPS_LAUNCH_OK("lanczos_h_staged_kernel");
      } else {
        lanczos_h_kernel<<<dim3((outW + 255) / 256, H, C), 256, 0, st>>>(in, h_out, db, dk, c.ksize, H, W, outW);
        IPS_LAUNCH_OK("lanczos_h_kernel");
      }
    }
    IPS_CUDA_OK(cudaGetLastError());
    v_in = h_out;
    v_W = outW;
  }
  if (need_v) {
    const LanczosCoeffs& c = cached_coeffs(H, outH);
    if (!uniform_v(c, v_in, out, C, H, v_W, outH, st)) {
      int* db = reinterpret_cast<int*>(base + L.vb);
      double* dk = reinterpret_cast<double*>(base + L.vk);
      IPS_CUDA_OK(cudaMemcpyAsync(db, c.bounds.data(), c.bounds.size() * sizeof(int), cudaMemcpyHostToDevice, st));
      IPS_CUDA_OK(cudaMemcpyAsync(dk, c.kk.data(), c.kk.size() * sizeof(double), cudaMemcpyHostToDevice, st));
      const int span_max = lanczos_span_max(c, outH, LZ_RV);
      const size_t smem = ((size_t)span_max * LZ_TX + (size_t)LZ_RV * c.ksize) * sizeof(double);
      if (smem <= LZ_SMEM_LIMIT) {
        IPS_CUDA_OK(cudaFuncSetAttribute(lanczos_v_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lanczos_v_staged_kernel<<<dim3((v_W + LZ_TX - 1) / LZ_TX, (outH + LZ_RV - 1) / LZ_RV, C), LZ_TX, smem, st>>>(
            v_in, out, db, dk, c.ksize, H, v_W, outH, span_max);
        IPS_LAUNCH_OK("lanczos_v_staged_kernel");
      } else {
        lanczos_v_kernel<<<dim3((v_W + 255) / 256, outH, C), 256, 0, st>>>(v_in, out, db, dk, c.ksize, H, v_W, outH);
        IPS_LAUNCH_OK("lanczos_v_kernel");
      }
    }
    IPS_CUDA_OK(cudaGetLastError());
  }
  return IPS_OK;
}
