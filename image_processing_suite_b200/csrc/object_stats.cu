// K3 -- per-object statistics over a label mask: a label-keyed segmented reduction.
//
// Replaces the CellProfiler MeasureObjectSizeShape / MeasureObjectIntensity subprocess
// (Feature_extraction_opt.py:166-167) and the regionprops conventions of
// Cellpose_GPU_s3fs.py:149-170: area, half-open bbox, centroid, and per channel
// sum (integrated intensity), mean, population std, min, max.
//
// Structure (three launches per batch of fields):
//   1. init      -- reset the per-(field, label) accumulator records.
//   2. scan      -- each thread owns an 8-pixel-wide column strip of one field and walks
//                   down R rows.  Per row it issues 128-bit loads of the labels and of every
//                   channel (+ illumination function) and folds the pixels into a register-
//                   resident run accumulator keyed by the current label.  Cellpose objects
//                   are compact, so a strip sees the same label for tens of rows; the
//                   accumulator is flushed to the label's global record only when the label
//                   changes or the strip ends (~9 flushes per cell instead of ~1000 pixel
//                   updates).  Flushes use integer atomics for area / bbox / coordinate sums
//                   (exact and order-independent) and 64-bit atomics for the intensity
//                   moments (float64 sums, or exact uint64 sums when there is no
//                   illumination function), so std has no fp32 cancellation problem.
//   3. compact   -- one block per field: prefix-sum over "area > 0" and emit dense rows in
//                   ascending label order, converting sums to mean / std / centroid.
// Warps are numbered field-fastest so that the warps reading the same piece of the
// plate-constant illumination function run together and share it through L1/L2.
#include "ips_common.cuh"

namespace ips {

constexpr int K3_THREADS = 128;
constexpr int K3_PX = 8;

// record layout in 8-byte words: [0] sum_y, [1] sum_x, [2] area | ymin, [3] xmin | ymax1,
// [4] xmax1 | pad, then per channel [5+3c] sum, [6+3c] sumsq, [7+3c] min | max.
__host__ __device__ constexpr int k3_record_words(int C) { return 5 + 3 * C; }

template <int C, bool HAS_ILLUM>
struct RunAcc {
  int label;
  unsigned area, sy, sx;
  int ymin, ylast, xmin, xmax;
  double fsum[HAS_ILLUM ? C : 1], fsq[HAS_ILLUM ? C : 1];
  unsigned isum[HAS_ILLUM ? 1 : C];
  unsigned long long isq[HAS_ILLUM ? 1 : C];
  unsigned mn[C], mx[C];  // float bits (non-negative floats order like unsigned) or integers

  __device__ __forceinline__ void reset(int l, int y) {
    label = l;
    area = sy = sx = 0u;
    ymin = ylast = y;
    xmin = 0x7fffffff;
    xmax = -1;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (HAS_ILLUM) { fsum[c] = 0.0; fsq[c] = 0.0; }
      else { isum[c] = 0u; isq[c] = 0ull; }
      mn[c] = 0xffffffffu;
      mx[c] = 0u;
    }
  }

  __device__ __forceinline__ void flush(unsigned long long* __restrict__ rec_base) const {
    if (label <= 0 || area == 0u) return;
    unsigned long long* r = rec_base + (size_t)(label - 1) * k3_record_words(C);
    unsigned* r32 = reinterpret_cast<unsigned*>(r);
    atomicAdd(r + 0, (unsigned long long)sy);
    atomicAdd(r + 1, (unsigned long long)sx);
    atomicAdd(r32 + 4, area);
    atomicMin(reinterpret_cast<int*>(r32 + 5), ymin);
    atomicMin(reinterpret_cast<int*>(r32 + 6), xmin);
    atomicMax(reinterpret_cast<int*>(r32 + 7), ylast + 1);
    atomicMax(reinterpret_cast<int*>(r32 + 8), xmax + 1);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (HAS_ILLUM) {
        atomicAdd(reinterpret_cast<double*>(r + 5 + 3 * c), fsum[c]);
        atomicAdd(reinterpret_cast<double*>(r + 6 + 3 * c), fsq[c]);
      } else {
        atomicAdd(r + 5 + 3 * c, (unsigned long long)isum[c]);
        atomicAdd(r + 6 + 3 * c, isq[c]);
      }
      atomicMin(r32 + 2 * (7 + 3 * c), mn[c]);
      atomicMax(r32 + 2 * (7 + 3 * c) + 1, mx[c]);
    }
  }
};

__global__ void object_stats_init_kernel(unsigned long long* __restrict__ rec, int* __restrict__ flags,
                                         size_t n_records, int words, int C, int F) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)F) flags[i] = 0;
  if (i >= n_records * (size_t)words) return;
  const int w = (int)(i % words);
  unsigned long long v = 0ull;
  if (w == 2) v = (unsigned long long)0x7fffffffu << 32;            // area = 0 | ymin = INT_MAX
  else if (w == 3) v = 0x7fffffffull;                               // xmin = INT_MAX | ymax1 = 0
  else if (w >= 5 && (w - 5) % 3 == 2) v = 0xffffffffull;           // min = UINT_MAX | max = 0
  (void)C;
  rec[i] = v;
}

template <int C, bool HAS_ILLUM, bool VEC, bool SKIP_BG>
__global__ void __launch_bounds__(K3_THREADS)
object_stats_scan_kernel(const int32_t* __restrict__ labels, const uint16_t* __restrict__ maxproj,
                         const float* __restrict__ illum, unsigned long long* __restrict__ rec,
                         int* __restrict__ flags, int Nmax, int F, int H, int W, int R,
                         int warps_x, long n_warps) {
  const long widx = (long)blockIdx.x * (K3_THREADS / 32) + (threadIdx.x >> 5);
  if (widx >= n_warps) return;
  const int lane = threadIdx.x & 31;
  const int f = (int)(widx % F);
  const long t = widx / F;
  const int wx = (int)(t % warps_x);
  const int band = (int)(t / warps_x);
  const int x0 = (wx * 32 + lane) * K3_PX;
  if (x0 >= W) return;
  const int y_begin = band * R;
  const int y_end = min(H, y_begin + R);
  const size_t plane = (size_t)H * W;
  const int32_t* lp = labels + (size_t)f * plane;
  const uint16_t* mp = maxproj + (size_t)f * C * plane;
  unsigned long long* rec_f = rec + (size_t)f * Nmax * k3_record_words(C);

  const uint64_t pol_stream = policy_evict_first();
  const uint64_t pol_keep = policy_evict_last();

  RunAcc<C, HAS_ILLUM> acc;
  acc.reset(0, y_begin);
  bool overflow = false;

  for (int y = y_begin; y < y_end; ++y) {
    const size_t off = (size_t)y * W + x0;
    int lab[K3_PX];
    uint4 px[C];
    uint4 il[HAS_ILLUM ? C : 1][2];
    if (VEC) {
      const uint4 l0 = ldg128_stream(lp + off, pol_stream);
      const uint4 l1 = ldg128_stream(lp + off + 4, pol_stream);
      lab[0] = (int)l0.x; lab[1] = (int)l0.y; lab[2] = (int)l0.z; lab[3] = (int)l0.w;
      lab[4] = (int)l1.x; lab[5] = (int)l1.y; lab[6] = (int)l1.z; lab[7] = (int)l1.w;
      if (SKIP_BG) {
        const int any = lab[0] | lab[1] | lab[2] | lab[3] | lab[4] | lab[5] | lab[6] | lab[7];
        if (any == 0) continue;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        px[c] = ldg128_stream(mp + (size_t)c * plane + off, pol_stream);
        if (HAS_ILLUM) {
          il[c][0] = ldg128_keep(illum + (size_t)c * plane + off, pol_keep);
          il[c][1] = ldg128_keep(illum + (size_t)c * plane + off + 4, pol_keep);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < K3_PX; ++i) lab[i] = (x0 + i < W) ? lp[off + i] : 0;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        unsigned short v[K3_PX];
        float d[K3_PX];
#pragma unroll
        for (int i = 0; i < K3_PX; ++i) {
          const bool in = x0 + i < W;
          v[i] = in ? mp[(size_t)c * plane + off + i] : (unsigned short)0;
          d[i] = (HAS_ILLUM && in) ? illum[(size_t)c * plane + off + i] : 1.0f;
        }
        px[c] = make_uint4(v[0] | ((unsigned)v[1] << 16), v[2] | ((unsigned)v[3] << 16),
                           v[4] | ((unsigned)v[5] << 16), v[6] | ((unsigned)v[7] << 16));
        if (HAS_ILLUM) {
          il[c][0] = make_uint4(__float_as_uint(d[0]), __float_as_uint(d[1]), __float_as_uint(d[2]), __float_as_uint(d[3]));
          il[c][1] = make_uint4(__float_as_uint(d[4]), __float_as_uint(d[5]), __float_as_uint(d[6]), __float_as_uint(d[7]));
        }
      }
    }
    if (!SKIP_BG || !VEC) {
      const int any = lab[0] | lab[1] | lab[2] | lab[3] | lab[4] | lab[5] | lab[6] | lab[7];
      if (any == 0) continue;
    }

    // values of this row, channel-major
    unsigned iv[C][K3_PX];
    float fv[HAS_ILLUM ? C : 1][K3_PX];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      unpack_u16x8(px[c], iv[c]);
      if (HAS_ILLUM) {
        const float d[8] = {__uint_as_float(il[c][0].x), __uint_as_float(il[c][0].y),
                            __uint_as_float(il[c][0].z), __uint_as_float(il[c][0].w),
                            __uint_as_float(il[c][1].x), __uint_as_float(il[c][1].y),
                            __uint_as_float(il[c][1].z), __uint_as_float(il[c][1].w)};
#pragma unroll
        for (int i = 0; i < K3_PX; ++i) fv[c][i] = __fdividef((float)iv[c][i], d[i]);
      }
    }

#pragma unroll
    for (int i = 0; i < K3_PX; ++i) {
      const int l = lab[i];
      if (l <= 0) continue;
      if (l > Nmax) { overflow = true; continue; }
      if (l != acc.label) {
        acc.flush(rec_f);
        acc.reset(l, y);
      }
      const int x = x0 + i;
      acc.area += 1u;
      acc.sy += (unsigned)y;
      acc.sx += (unsigned)x;
      acc.ylast = y;
      acc.xmin = min(acc.xmin, x);
      acc.xmax = max(acc.xmax, x);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if (HAS_ILLUM) {
          const float v = fv[c][i];
          const double dv = (double)v;
          acc.fsum[c] += dv;
          acc.fsq[c] = fma(dv, dv, acc.fsq[c]);
          const unsigned b = __float_as_uint(v);
          acc.mn[c] = min(acc.mn[c], b);
          acc.mx[c] = max(acc.mx[c], b);
        } else {
          const unsigned v = iv[c][i];
          acc.isum[c] += v;
          acc.isq[c] += (unsigned long long)(v * v);
          acc.mn[c] = min(acc.mn[c], v);
          acc.mx[c] = max(acc.mx[c], v);
        }
      }
    }
  }
  acc.flush(rec_f);
  if (overflow) atomicOr(flags + f, 1);
}

// One block per field: dense rows in ascending label order.
template <bool HAS_ILLUM>
__global__ void __launch_bounds__(1024)
object_stats_compact_kernel(const unsigned long long* __restrict__ rec, const int* __restrict__ flags,
                            int32_t* __restrict__ n_objects, int32_t* __restrict__ ints,
                            float* __restrict__ flts, int Nmax, int C, float intensity_scale) {
  const int f = blockIdx.x;
  const int words = k3_record_words(C);
  const int nf = 2 + 5 * C;
  const unsigned long long* rec_f = rec + (size_t)f * Nmax * words;
  int32_t* ints_f = ints + (size_t)f * Nmax * 6;
  float* flts_f = flts + (size_t)f * Nmax * nf;
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double sc = (double)intensity_scale;
  for (int start = 0; start < Nmax; start += blockDim.x) {
    const int li = start + threadIdx.x;  // label - 1
    const unsigned* r32 = nullptr;
    unsigned area = 0;
    if (li < Nmax) {
      r32 = reinterpret_cast<const unsigned*>(rec_f + (size_t)li * words);
      area = r32[4];
    }
    const bool present = area > 0u;
    const unsigned bal = __ballot_sync(0xffffffffu, present);
    const int wpre = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int wbase = 0, total = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) {
      const int tcount = warp_tot[w];
      if (w < warp) wbase += tcount;
      total += tcount;
    }
    const int base = base_s;
    if (present) {
      const int row = base + wbase + wpre;
      const unsigned long long* r = rec_f + (size_t)li * words;
      int32_t* o = ints_f + (size_t)row * 6;
      o[0] = li + 1;
      o[1] = (int)area;
      o[2] = (int)r32[5];
      o[3] = (int)r32[6];
      o[4] = (int)r32[7];
      o[5] = (int)r32[8];
      float* q = flts_f + (size_t)row * nf;
      const double n = (double)area;
      q[0] = (float)((double)r[0] / n);
      q[1] = (float)((double)r[1] / n);
      for (int c = 0; c < C; ++c) {
        double s, mean, var;
        float mn, mx;
        if (HAS_ILLUM) {
          s = __longlong_as_double((long long)r[5 + 3 * c]);
          const double sq = __longlong_as_double((long long)r[6 + 3 * c]);
          mean = s / n;
          var = sq / n - mean * mean;
          mn = __uint_as_float(r32[2 * (7 + 3 * c)]);
          mx = __uint_as_float(r32[2 * (7 + 3 * c) + 1]);
        } else {
          const unsigned long long is = r[5 + 3 * c], isq = r[6 + 3 * c];
          s = (double)is;
          mean = s / n;
          // exact integer numerator n*sum(x^2) - (sum x)^2, so constant objects give 0
          const unsigned __int128 num = (unsigned __int128)isq * area - (unsigned __int128)is * is;
          var = (double)(unsigned long long)(num >> 32) * 4294967296.0 + (double)(unsigned)(num & 0xffffffffu);
          var = var / (n * n);
          mn = (float)r32[2 * (7 + 3 * c)];
          mx = (float)r32[2 * (7 + 3 * c) + 1];
        }
        if (!(var > 0.0)) var = 0.0;
        float* qc = q + 2 + 5 * c;
        qc[0] = (float)(s * sc);
        qc[1] = (float)(mean * sc);
        qc[2] = (float)(sqrt(var) * sc);
        qc[3] = (float)((double)mn * sc);
        qc[4] = (float)((double)mx * sc);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s = base + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_objects[f] = flags[f] ? -1 : base_s;
}

template <int C, bool HAS_ILLUM>
static void launch_scan(bool vec, bool skip_bg, int blocks, cudaStream_t st, const int32_t* labels,
                        const uint16_t* maxproj, const float* illum, unsigned long long* rec,
                        int* flags, int Nmax, int F, int H, int W, int R, int warps_x, long n_warps) {
  if (vec) {
    if (skip_bg)
      object_stats_scan_kernel<C, HAS_ILLUM, true, true><<<blocks, K3_THREADS, 0, st>>>(labels, maxproj, illum, rec, flags, Nmax, F, H, W, R, warps_x, n_warps);
    else
      object_stats_scan_kernel<C, HAS_ILLUM, true, false><<<blocks, K3_THREADS, 0, st>>>(labels, maxproj, illum, rec, flags, Nmax, F, H, W, R, warps_x, n_warps);
  } else {
    object_stats_scan_kernel<C, HAS_ILLUM, false, false><<<blocks, K3_THREADS, 0, st>>>(labels, maxproj, illum, rec, flags, Nmax, F, H, W, R, warps_x, n_warps);
  }
}

template <int C>
static void launch_scan_c(bool has_illum, bool vec, bool skip_bg, int blocks, cudaStream_t st,
                          const int32_t* labels, const uint16_t* maxproj, const float* illum,
                          unsigned long long* rec, int* flags, int Nmax, int F, int H, int W, int R,
                          int warps_x, long n_warps) {
  if (has_illum) launch_scan<C, true>(vec, skip_bg, blocks, st, labels, maxproj, illum, rec, flags, Nmax, F, H, W, R, warps_x, n_warps);
  else launch_scan<C, false>(vec, skip_bg, blocks, st, labels, maxproj, illum, rec, flags, Nmax, F, H, W, R, warps_x, n_warps);
}

static size_t k3_records_bytes(int F, int C, int Nmax) {
  return round_up((size_t)F * Nmax * k3_record_words(C) * 8, 256);
}

// tuning knobs (read once): IPS_K3_ROWS = rows per strip, IPS_K3_SKIP_BG = 0/1
static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

}  // namespace ips

using namespace ips;

extern "C" size_t ips_object_stats_workspace_bytes(int F, int C, int Nmax) {
  if (F <= 0 || C <= 0 || Nmax <= 0) return 0;
  return k3_records_bytes(F, C, Nmax) + round_up((size_t)F * sizeof(int), 256);
}

extern "C" int ips_object_stats(const int32_t* labels, const uint16_t* maxproj, const float* illum,
                                float intensity_scale, int32_t* n_objects, int32_t* ints,
                                float* flts, int Nmax, void* ws, size_t ws_bytes, int F, int C,
                                int H, int W, ips_stream_t stream) {
  if (!labels || !maxproj || !n_objects || !ints || !flts)
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_object_stats: NULL pointer argument");
  if (F < 0 || H <= 0 || W <= 0 || Nmax <= 0)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_object_stats: bad shape F=%d H=%d W=%d Nmax=%d", F, H, W, Nmax);
  if (C < 1 || C > 8) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_object_stats: C must be in 1..8 (got %d)", C);
  if (F == 0) return IPS_OK;
  const size_t need = ips_object_stats_workspace_bytes(F, C, Nmax);
  if (ws == nullptr || ws_bytes < need)
    IPS_FAIL(IPS_ERR_NOMEM, "ips_object_stats: needs %zu workspace bytes (got %zu)", need, ws_bytes);
  if (!aligned16(ws)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_object_stats: workspace not 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned long long* rec = reinterpret_cast<unsigned long long*>(ws);
  int* flags = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + k3_records_bytes(F, C, Nmax));
  const int words = k3_record_words(C);
  const size_t n_records = (size_t)F * Nmax;
  {
    const size_t n = n_records * words;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    object_stats_init_kernel<<<blocks, 256, 0, st>>>(rec, flags, n_records, words, C, F);
    IPS_LAUNCH_OK("object_stats_init_kernel");
  }
  const bool vec = (W % 8 == 0) && aligned16(labels) && aligned16(maxproj) && aligned16(illum);
  static const int env_rows = env_int("IPS_K3_ROWS", 0);
  static const int env_skip = env_int("IPS_K3_SKIP_BG", 0);
  const int warps_x = (W + 32 * K3_PX - 1) / (32 * K3_PX);
  // rows per strip: long strips mean fewer flushes; short strips mean more warps in flight.
  int R = env_rows > 0 ? env_rows : 64;
  if (env_rows <= 0) {
    const long target_warps = (long)sm_count() * 32;
    while (R > 8 && (long)((H + R - 1) / R) * warps_x * F < target_warps) R >>= 1;
  }
  const int bands = (H + R - 1) / R;
  const long n_warps = (long)bands * warps_x * F;
  const long blocks_l = (n_warps + (K3_THREADS / 32) - 1) / (K3_THREADS / 32);
  if (blocks_l > 0x7fffffffL) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_object_stats: batch too large for one launch");
  const int blocks = (int)blocks_l;
  const bool has_illum = illum != nullptr;
  const bool skip_bg = env_skip != 0;
#define IPS_K3_CASE(CC) \
  case CC: launch_scan_c<CC>(has_illum, vec, skip_bg, blocks, st, labels, maxproj, illum, rec, flags, Nmax, F, H, W, R, warps_x, n_warps); break
  switch (C) {
    IPS_K3_CASE(1); IPS_K3_CASE(2); IPS_K3_CASE(3); IPS_K3_CASE(4);
    IPS_K3_CASE(5); IPS_K3_CASE(6); IPS_K3_CASE(7); IPS_K3_CASE(8);
  }
#undef IPS_K3_CASE
  IPS_LAUNCH_OK("object_stats_scan_kernel");
  if (has_illum)
    object_stats_compact_kernel<true><<<F, 1024, 0, st>>>(rec, flags, n_objects, ints, flts, Nmax, C, intensity_scale);
  else
    object_stats_compact_kernel<false><<<F, 1024, 0, st>>>(rec, flags, n_objects, ints, flts, Nmax, C, intensity_scale);
  IPS_LAUNCH_OK("object_stats_compact_kernel");
  return IPS_OK;
}
