// K1 -- fused z-max projection -> illumination divide -> b x b sum binning (+ optional
// PercentMaximal side reduction).  One streaming pass: every input byte is read once and
// every output byte written once.
//
// Replaces np.maximum.reduce (MaxProjection.py:45), img.astype(float)/illum
// (Illumination_QC_mult.py:145-150, Cellpose_GPU_s3fs.py:72), PercentMaximal
// (Illumination_QC_mult.py:73-95) and north_star's 2x2/4x4 sum re-binning.
//
// Work decomposition: one work item = BIN rows x 8 columns of one (field, channel) plane,
// i.e. BIN x Z 128-bit loads of raw data, BIN x 2 128-bit loads of the illumination
// function, BIN 128-bit stores of the max projection and one 8/16/32-byte store of the
// binned row.  Items of a plane are flattened so consecutive lanes touch consecutive
// 16-byte words (fully coalesced, rows are multiples of 16 bytes).  Blocks are numbered
// field-fastest: the F blocks that need the same piece of the (plate-constant)
// illumination function are adjacent in launch order, so it is fetched from HBM once per
// launch and served from L2 (evict_last policy) to the other F-1 fields, while raw data
// and outputs stream through with evict_first.
#include <stdlib.h>

#include "ips_common.cuh"

namespace ips {

// preprocess_tma.cu: the TMA-staged variant of this pass (IPS_K1_TMA=1 selects it)
int preprocess_tma_try(const uint16_t* raw, const float* illum, uint16_t* maxproj, void* binned, int bin, int F,
                       int C, int Z, int H, int W, cudaStream_t st);
static bool k1_use_tma() {
  static const bool v = [] {
    const char* s = getenv("IPS_K1_TMA");
    return s && *s && atoi(s) != 0;
  }();
  return v;
}

#define PCT_NEG_INF (__longlong_as_double(0xfff0000000000000ll))

struct PctPartial {
  double maxv;
  unsigned long long cnt;
};

__device__ __forceinline__ void pct_merge(double& m, unsigned long long& n, double om,
                                          unsigned long long on) {
  if (om > m) {
    m = om;
    n = on;
  } else if (om == m) {
    n += on;
  }
}

template <int THREADS>
__device__ __forceinline__ void pct_block_reduce(double m, unsigned long long n, PctPartial* out) {
  __shared__ double s_m[THREADS / 32];
  __shared__ unsigned long long s_n[THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double om = __shfl_xor_sync(0xffffffffu, m, o);
    unsigned long long on = __shfl_xor_sync(0xffffffffu, n, o);
    pct_merge(m, n, om, on);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_m[warp] = m;
    s_n[warp] = n;
  }
  __syncthreads();
  if (warp == 0) {
    m = lane < THREADS / 32 ? s_m[lane] : PCT_NEG_INF;
    n = lane < THREADS / 32 ? s_n[lane] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double om = __shfl_xor_sync(0xffffffffu, m, o);
      unsigned long long on = __shfl_xor_sync(0xffffffffu, n, o);
      pct_merge(m, n, om, on);
    }
    if (lane == 0) {
      out->maxv = m;
      out->cnt = n;
    }
  }
}

constexpr int K1_THREADS = 256;

// ZT > 0: compile-time number of planes (all loads issued up front); ZT == 0: runtime Z.
template <int BIN, int ZT, bool HAS_ILLUM, bool PCT>
__global__ void __launch_bounds__(K1_THREADS)
preprocess_vec_kernel(const uint16_t* __restrict__ raw, const float* __restrict__ illum,
                      uint16_t* __restrict__ maxproj, float* __restrict__ corrected,
                      void* __restrict__ binned, PctPartial* __restrict__ pct_ws, int F, int C,
                      int Z, int H, int W, int n_chunks) {
  const int bid = blockIdx.x;
  const int f = bid % F;
  const int t = bid / F;
  const int chunk = t % n_chunks;
  const int c = t / n_chunks;
  const int G = W >> 3;                 // 16-byte groups per row
  const int n_items = (H / BIN) * G;
  const int item = chunk * K1_THREADS + threadIdx.x;
  const bool active = item < n_items;

  const uint64_t pol_stream = policy_evict_first();
  const uint64_t pol_keep = policy_evict_last();

  double pmax = PCT_NEG_INF;
  unsigned long long pcnt = 0;

  if (active) {
    const int rb = item / G;
    const int g = item - rb * G;
    const int y0 = rb * BIN;
    const int x0 = g << 3;
    const size_t plane = (size_t)H * W;
    const size_t fc = (size_t)f * C + c;
    const int nz = ZT > 0 ? ZT : Z;
    const uint16_t* rp = raw + fc * nz * plane + (size_t)y0 * W + x0;

    // ---- phase 1: issue every load of this item -------------------------------------
    uint4 m[BIN];
    if (ZT > 0) {
      uint4 v[BIN][ZT > 0 ? ZT : 1];
#pragma unroll
      for (int r = 0; r < BIN; ++r)
#pragma unroll
        for (int z = 0; z < ZT; ++z)
          v[r][z] = ldg128_stream(rp + (size_t)z * plane + (size_t)r * W, pol_stream);
#pragma unroll
      for (int r = 0; r < BIN; ++r) {
        m[r] = v[r][0];
#pragma unroll
        for (int z = 1; z < ZT; ++z) m[r] = vmax_u16x8(m[r], v[r][z]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < BIN; ++r) m[r] = ldg128_stream(rp + (size_t)r * W, pol_stream);
      for (int z = 1; z < nz; ++z) {
#pragma unroll
        for (int r = 0; r < BIN; ++r)
          m[r] = vmax_u16x8(m[r], ldg128_stream(rp + (size_t)z * plane + (size_t)r * W, pol_stream));
      }
    }
    uint4 il[BIN][2];
    if (HAS_ILLUM) {
      const float* ip = illum + (size_t)c * plane + (size_t)y0 * W + x0;
#pragma unroll
      for (int r = 0; r < BIN; ++r) {
        il[r][0] = ldg128_keep(ip + (size_t)r * W, pol_keep);
        il[r][1] = ldg128_keep(ip + (size_t)r * W + 4, pol_keep);
      }
    }

    // ---- phase 2: store max projection, divide, bin ------------------------------------
    if (maxproj != nullptr) {
      uint16_t* mp = maxproj + fc * plane + (size_t)y0 * W + x0;
#pragma unroll
      for (int r = 0; r < BIN; ++r) stg128_stream(mp + (size_t)r * W, m[r], pol_stream);
    }

    constexpr int NB = 8 / BIN;  // binned outputs per item
    if (HAS_ILLUM) {
      float bs[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) bs[j] = 0.f;
#pragma unroll
      for (int r = 0; r < BIN; ++r) {
        float x[8], q[8];
        unpack_u16x8(m[r], x);
        const float d[8] = {__uint_as_float(il[r][0].x), __uint_as_float(il[r][0].y),
                            __uint_as_float(il[r][0].z), __uint_as_float(il[r][0].w),
                            __uint_as_float(il[r][1].x), __uint_as_float(il[r][1].y),
                            __uint_as_float(il[r][1].z), __uint_as_float(il[r][1].w)};
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = fast_div(x[i], d[i]);
        if (corrected != nullptr) {
          float* cp = corrected + fc * plane + (size_t)(y0 + r) * W + x0;
          stg128_stream(cp, make_uint4(__float_as_uint(q[0]), __float_as_uint(q[1]),
                                       __float_as_uint(q[2]), __float_as_uint(q[3])), pol_stream);
          stg128_stream(cp + 4, make_uint4(__float_as_uint(q[4]), __float_as_uint(q[5]),
                                           __float_as_uint(q[6]), __float_as_uint(q[7])), pol_stream);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) bs[i / BIN] += q[i];
        if (PCT) {
          // Exact float64 PercentMaximal: the fp64 maximum can only sit at pixels whose
          // fp32 quotient is within rounding distance of the thread's fp32 maximum, so
          // only those few pay for a double-precision divide.
          float qm = q[0];
#pragma unroll
          for (int i = 1; i < 8; ++i) qm = fmaxf(qm, q[i]);
          const float thr = qm * (1.0f - 1e-6f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (q[i] >= thr) {
              const double qd = (double)x[i] / (double)d[i];
              pct_merge(pmax, pcnt, qd, 1ull);
            }
          }
        }
      }
      if (binned != nullptr) {
        float* bp = reinterpret_cast<float*>(binned) + fc * (plane / (BIN * BIN)) +
                    (size_t)rb * (W / BIN) + (size_t)g * NB;
        if (NB == 8) {
          stg128_stream(bp, make_uint4(__float_as_uint(bs[0]), __float_as_uint(bs[1]),
                                       __float_as_uint(bs[2]), __float_as_uint(bs[3])), pol_stream);
          stg128_stream(bp + 4, make_uint4(__float_as_uint(bs[4 % NB]), __float_as_uint(bs[5 % NB]),
                                           __float_as_uint(bs[6 % NB]), __float_as_uint(bs[7 % NB])), pol_stream);
        } else if (NB == 4) {
          stg128_stream(bp, make_uint4(__float_as_uint(bs[0]), __float_as_uint(bs[1 % NB]),
                                       __float_as_uint(bs[2 % NB]), __float_as_uint(bs[3 % NB])), pol_stream);
        } else {
          stg64_stream(bp, make_uint2(__float_as_uint(bs[0]), __float_as_uint(bs[1 % NB])), pol_stream);
        }
      }
    } else {
      uint32_t bs[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) bs[j] = 0u;
#pragma unroll
      for (int r = 0; r < BIN; ++r) {
        uint32_t x[8];
        unpack_u16x8(m[r], x);
#pragma unroll
        for (int i = 0; i < 8; ++i) bs[i / BIN] += x[i];
        if (PCT) {
          uint32_t xm = x[0];
#pragma unroll
          for (int i = 1; i < 8; ++i) xm = max(xm, x[i]);
          unsigned long long k = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) k += (x[i] == xm);
          pct_merge(pmax, pcnt, (double)xm, k);
        }
      }
      if (binned != nullptr) {
        uint32_t* bp = reinterpret_cast<uint32_t*>(binned) + fc * (plane / (BIN * BIN)) +
                       (size_t)rb * (W / BIN) + (size_t)g * NB;
        if (NB == 8) {
          stg128_stream(bp, make_uint4(bs[0], bs[1 % NB], bs[2 % NB], bs[3 % NB]), pol_stream);
          stg128_stream(bp + 4, make_uint4(bs[4 % NB], bs[5 % NB], bs[6 % NB], bs[7 % NB]), pol_stream);
        } else if (NB == 4) {
          stg128_stream(bp, make_uint4(bs[0], bs[1 % NB], bs[2 % NB], bs[3 % NB]), pol_stream);
        } else {
          stg64_stream(bp, make_uint2(bs[0], bs[1 % NB]), pol_stream);
        }
      }
    }
  }

  if (PCT) pct_block_reduce<K1_THREADS>(pmax, pcnt, pct_ws + ((size_t)f * C + c) * n_chunks + chunk);
}

// Any width / alignment: one thread per binned cell, scalar accesses.  Correctness path
// for shapes the vector kernel does not take (W % 8 != 0 or unaligned bases).
template <bool HAS_ILLUM, bool PCT>
__global__ void __launch_bounds__(K1_THREADS)
preprocess_scalar_kernel(const uint16_t* __restrict__ raw, const float* __restrict__ illum,
                         uint16_t* __restrict__ maxproj, float* __restrict__ corrected,
                         void* __restrict__ binned, PctPartial* __restrict__ pct_ws, int bin, int F,
                         int C, int Z, int H, int W, int n_chunks) {
  const int bid = blockIdx.x;
  const int f = bid % F;
  const int t = bid / F;
  const int chunk = t % n_chunks;
  const int c = t / n_chunks;
  const int Wb = W / bin, Hb = H / bin;
  const int cell = chunk * K1_THREADS + threadIdx.x;
  double pmax = PCT_NEG_INF;
  unsigned long long pcnt = 0;
  if (cell < Wb * Hb) {
    const int by = cell / Wb, bx = cell - by * Wb;
    const size_t plane = (size_t)H * W;
    const size_t fc = (size_t)f * C + c;
    float fs = 0.f;
    uint32_t is = 0;
    for (int dy = 0; dy < bin; ++dy)
      for (int dx = 0; dx < bin; ++dx) {
        const size_t off = (size_t)(by * bin + dy) * W + (bx * bin + dx);
        uint32_t m = 0;
        for (int z = 0; z < Z; ++z) m = max(m, (uint32_t)raw[(fc * Z + z) * plane + off]);
        if (maxproj != nullptr) maxproj[fc * plane + off] = (uint16_t)m;
        if (HAS_ILLUM) {
          const float d = illum[(size_t)c * plane + off];
          const float q = fast_div((float)m, d);
          if (corrected != nullptr) corrected[fc * plane + off] = q;
          fs += q;
          if (PCT) pct_merge(pmax, pcnt, (double)m / (double)d, 1ull);
        } else {
          is += m;
          if (PCT) pct_merge(pmax, pcnt, (double)m, 1ull);
        }
      }
    if (binned != nullptr) {
      const size_t o = fc * ((size_t)Hb * Wb) + (size_t)by * Wb + bx;
      if (HAS_ILLUM) reinterpret_cast<float*>(binned)[o] = fs;
      else reinterpret_cast<uint32_t*>(binned)[o] = is;
    }
  }
  if (PCT) pct_block_reduce<K1_THREADS>(pmax, pcnt, pct_ws + ((size_t)f * C + c) * n_chunks + chunk);
}

__global__ void pct_finalize_kernel(const PctPartial* __restrict__ ws, double* __restrict__ out,
                                    int n_chunks, double n_pixels) {
  const int fc = blockIdx.x;
  double m = PCT_NEG_INF;
  unsigned long long n = 0;
  for (int i = threadIdx.x; i < n_chunks; i += blockDim.x) {
    const PctPartial p = ws[(size_t)fc * n_chunks + i];
    pct_merge(m, n, p.maxv, p.cnt);
  }
  __shared__ PctPartial res;
  pct_block_reduce<256>(m, n, &res);
  __syncthreads();
  if (threadIdx.x == 0) out[fc] = 100.0 * (double)res.cnt / n_pixels;
}

template <int BIN, bool HAS_ILLUM, bool PCT>
static void launch_vec(int Z, int grid, cudaStream_t st, const uint16_t* raw, const float* illum,
                       uint16_t* maxproj, float* corrected, void* binned, PctPartial* pw, int F,
                       int C, int H, int W, int n_chunks) {
#define IPS_K1_CASE(ZT)                                                                   \
  preprocess_vec_kernel<BIN, ZT, HAS_ILLUM, PCT><<<grid, K1_THREADS, 0, st>>>(            \
      raw, illum, maxproj, corrected, binned, pw, F, C, Z, H, W, n_chunks)
  switch (Z) {
    case 1: IPS_K1_CASE(1); break;
    case 2: IPS_K1_CASE(2); break;
    case 3: IPS_K1_CASE(3); break;
    case 4: IPS_K1_CASE(4); break;
    case 5: IPS_K1_CASE(5); break;
    default: IPS_K1_CASE(0); break;
  }
#undef IPS_K1_CASE
}

template <bool HAS_ILLUM, bool PCT>
static void launch_vec_bin(int bin, int Z, int grid, cudaStream_t st, const uint16_t* raw,
                           const float* illum, uint16_t* maxproj, float* corrected, void* binned,
                           PctPartial* pw, int F, int C, int H, int W, int n_chunks) {
  if (bin == 1) launch_vec<1, HAS_ILLUM, PCT>(Z, grid, st, raw, illum, maxproj, corrected, binned, pw, F, C, H, W, n_chunks);
  else if (bin == 2) launch_vec<2, HAS_ILLUM, PCT>(Z, grid, st, raw, illum, maxproj, corrected, binned, pw, F, C, H, W, n_chunks);
  else launch_vec<4, HAS_ILLUM, PCT>(Z, grid, st, raw, illum, maxproj, corrected, binned, pw, F, C, H, W, n_chunks);
}

static bool k1_vector_ok(const void* raw, const void* illum, const void* maxproj,
                         const void* corrected, const void* binned, int W, int bin) {
  if (W % 8) return false;
  if (!aligned16(raw) || !aligned16(illum) || !aligned16(maxproj) || !aligned16(corrected)) return false;
  if (binned && (reinterpret_cast<uintptr_t>(binned) & (bin == 4 ? 7u : 15u))) return false;
  return true;
}

static int k1_chunks(int H, int W, int bin, bool vec) {
  const long items = vec ? (long)(H / bin) * (W / 8) : (long)(H / bin) * (W / bin);
  return (int)((items + K1_THREADS - 1) / K1_THREADS);
}

}  // namespace ips

using namespace ips;

extern "C" size_t ips_preprocess_workspace_bytes(int F, int C, int H, int W, int bin) {
  if (F <= 0 || C <= 0 || H <= 0 || W <= 0 || (bin != 1 && bin != 2 && bin != 4)) return 0;
  // sized for the scalar decomposition (the larger of the two)
  const int n_chunks = k1_chunks(H, W, bin, false);
  return round_up((size_t)F * C * n_chunks * sizeof(PctPartial), 256);
}

extern "C" int ips_preprocess_fused(const uint16_t* raw, const float* illum, uint16_t* maxproj,
                                    float* corrected, void* binned, int bin, double* pct_maximal,
                                    void* ws, size_t ws_bytes, int F, int C, int Z, int H, int W,
                                    ips_stream_t stream) {
  if (raw == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_preprocess_fused: raw is NULL");
  if (F < 0 || C <= 0 || Z <= 0 || H <= 0 || W <= 0)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_preprocess_fused: bad shape F=%d C=%d Z=%d H=%d W=%d", F, C, Z, H, W);
  if (bin != 1 && bin != 2 && bin != 4)
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_preprocess_fused: bin must be 1, 2 or 4 (got %d)", bin);
  if (H % bin || W % bin)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_preprocess_fused: %dx%d not divisible by bin %d", H, W, bin);
  if (corrected != nullptr && illum == nullptr)
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_preprocess_fused: corrected output requires illum");
  if (F == 0) return IPS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = k1_vector_ok(raw, illum, maxproj, corrected, binned, W, bin);
  const int n_chunks = k1_chunks(H, W, bin, vec);
  const long grid_l = (long)n_chunks * C * F;
  if (grid_l > 0x7fffffffL) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_preprocess_fused: batch too large for one launch");
  const int grid = (int)grid_l;
  const bool pct = pct_maximal != nullptr;
  PctPartial* pw = reinterpret_cast<PctPartial*>(ws);
  if (pct) {
    const size_t need = (size_t)F * C * n_chunks * sizeof(PctPartial);
    if (ws == nullptr || ws_bytes < need)
      IPS_FAIL(IPS_ERR_NOMEM, "ips_preprocess_fused: pct_maximal needs %zu workspace bytes (got %zu)", need, ws_bytes);
    if (!aligned16(ws)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_preprocess_fused: workspace not 16-byte aligned");
  }
  if (vec && !pct && corrected == nullptr && k1_use_tma()) {
    const int rc = preprocess_tma_try(raw, illum, maxproj, binned, bin, F, C, Z, H, W, st);
    if (rc <= 0) return rc;          // launched (0) or failed (< 0); 1 = shape not supported, fall through
  }
  if (vec) {
    if (illum != nullptr) {
      if (pct) launch_vec_bin<true, true>(bin, Z, grid, st, raw, illum, maxproj, corrected, binned, pw, F, C, H, W, n_chunks);
      else launch_vec_bin<true, false>(bin, Z, grid, st, raw, illum, maxproj, corrected, binned, pw, F, C, H, W, n_chunks);
    } else {
      if (pct) launch_vec_bin<false, true>(bin, Z, grid, st, raw, illum, maxproj, corrected, binned, pw, F, C, H, W, n_chunks);
      else launch_vec_bin<false, false>(bin, Z, grid, st, raw, illum, maxproj, corrected, binned, pw, F, C, H, W, n_chunks);
    }
    IPS_LAUNCH_OK("preprocess_vec_kernel");
  } else {
    if (illum != nullptr) {
      if (pct) preprocess_scalar_kernel<true, true><<<grid, K1_THREADS, 0, st>>>(raw, illum, maxproj, corrected, binned, pw, bin, F, C, Z, H, W, n_chunks);
      else preprocess_scalar_kernel<true, false><<<grid, K1_THREADS, 0, st>>>(raw, illum, maxproj, corrected, binned, pw, bin, F, C, Z, H, W, n_chunks);
    } else {
      if (pct) preprocess_scalar_kernel<false, true><<<grid, K1_THREADS, 0, st>>>(raw, illum, maxproj, corrected, binned, pw, bin, F, C, Z, H, W, n_chunks);
      else preprocess_scalar_kernel<false, false><<<grid, K1_THREADS, 0, st>>>(raw, illum, maxproj, corrected, binned, pw, bin, F, C, Z, H, W, n_chunks);
    }
    IPS_LAUNCH_OK("preprocess_scalar_kernel");
  }
  if (pct) {
    pct_finalize_kernel<<<F * C, 256, 0, st>>>(pw, pct_maximal, n_chunks, (double)H * (double)W);
    IPS_LAUNCH_OK("pct_finalize_kernel");
  }
  return IPS_OK;
}
