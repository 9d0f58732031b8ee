// K1 + K3 in one pass: z-max projection -> illumination divide -> b x b sum binning AND the
// per-object statistics over the label mask, without writing the max projection out and
// reading it (and the illumination function) back.  Every input byte of a field -- raw
// z-stack, label mask -- is read exactly once; the plate-constant illumination function is
// served from L2 to all fields of the launch.
//
// Replaces np.maximum.reduce (MaxProjection.py:45), img.astype(float)/illum
// (Illumination_QC_mult.py:145-150, Cellpose_GPU_s3fs.py:72), north_star's sum re-binning and
// the CellProfiler MeasureObject* subprocess (Feature_extraction_opt.py:166-167).
//
// Decomposition (object_accum.cuh): a CTA of 4 warps owns a 256-column x 4*BIN-row tile; a
// lane owns BIN rows x 8 columns, i.e. one 128-bit word per row of every plane, and walks
// the channels: Z x BIN 128-bit loads of raw data and 2 x BIN of the illumination function,
// packed-uint16 max, store of the max projection, divide, bin, store of the binned row, then
// the lane -> warp -> CTA -> global label-keyed reduction of that channel's moments.
#include <stdlib.h>

#include "object_accum.cuh"

namespace ips {

template <int BIN, int ZT, bool HAS_ILLUM, int MINB>
__global__ void __launch_bounds__(OA_THREADS, MINB)
field_fused_kernel(const uint16_t* __restrict__ raw, const float* __restrict__ illum,
                   const int32_t* __restrict__ labels, uint16_t* __restrict__ maxproj,
                   void* __restrict__ binned, unsigned long long* __restrict__ rec, int* __restrict__ flags,
                   int Nmax, int F, int C, int Z, int H, int W, int tiles_x) {
  __shared__ OaShared sh;
  oa_init_shared(sh);
  const int bid = blockIdx.x;
  const int f = bid % F;
  const int t = bid / F;
  const int tile_x = t % tiles_x, tile_y = t / tiles_x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = tile_y * OA_WARPS + warp;      // binned row / row block
  const int y0 = rb * BIN;
  const int g = tile_x * 32 + lane;             // 16-byte group along the row
  const int x0 = g * OA_PX;
  const size_t plane = (size_t)H * W;
  const bool active = x0 < W && y0 < H;         // H % BIN == 0: a row block is all in or all out
  const int nz = ZT > 0 ? ZT : Z;
  const uint64_t pol_stream = policy_evict_first();
  const uint64_t pol_keep = policy_evict_last();
  unsigned long long* rec_f = rec + (size_t)f * Nmax * k3_record_words(C);
  const int32_t* lp = labels + (size_t)f * plane + (size_t)y0 * W + x0;

  int lab[BIN][OA_PX];
#pragma unroll
  for (int r = 0; r < BIN; ++r) {
    if (active) {
      const uint4 l0 = ldg128_stream(lp + (size_t)r * W, pol_stream);
      const uint4 l1 = ldg128_stream(lp + (size_t)r * W + 4, pol_stream);
      lab[r][0] = (int)l0.x; lab[r][1] = (int)l0.y; lab[r][2] = (int)l0.z; lab[r][3] = (int)l0.w;
      lab[r][4] = (int)l1.x; lab[r][5] = (int)l1.y; lab[r][6] = (int)l1.z; lab[r][7] = (int)l1.w;
    } else {
#pragma unroll
      for (int i = 0; i < OA_PX; ++i) lab[r][i] = 0;
    }
  }
  // Pull channel c's rows towards L2 one channel ahead of their use: at 64 registers there is
  // no room for a second register set, but a prefetch needs none.  One lane in four (u16
  // rows: 64 bytes apart) / in two (float rows) touches every 128-byte line of the warp's span.
  auto prefetch_channel = [&](int c) {
    if (!active) return;
    if ((lane & 3) == 0) {
      const uint16_t* rp = raw + ((size_t)f * C + c) * nz * plane + (size_t)y0 * W + x0;
      for (int z = 0; z < nz; ++z)
#pragma unroll
        for (int r = 0; r < BIN; ++r)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (size_t)z * plane + (size_t)r * W));
    }
    if (HAS_ILLUM && (lane & 1) == 0) {
      const float* ip = illum + (size_t)c * plane + (size_t)y0 * W + x0;
#pragma unroll
      for (int r = 0; r < BIN; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(ip + (size_t)r * W));
    }
  };
  prefetch_channel(0);

  bool overflow = false;
  OaLane<BIN> L;
  oa_begin<BIN>(L, lab, lp, W, Nmax, y0, x0, sh, rec_f, C, overflow);
  const bool warp_fg = __any_sync(OA_FULL, (L.m1 | L.m2 | L.m3) != 0u);

  for (int c = 0; c < C; ++c) {
    const size_t fc = (size_t)f * C + c;
    uint4 m[BIN];
    uint4 il[BIN][2];
#pragma unroll
    for (int r = 0; r < BIN; ++r) {
      m[r] = make_uint4(0u, 0u, 0u, 0u);
      il[r][0] = il[r][1] = make_uint4(0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u);
    }
    if (active) {
      const uint16_t* rp = raw + fc * nz * plane + (size_t)y0 * W + x0;
      // next channel towards L2 (distance 1 measured best: 0.744 ms per 16 fields against 0.767
      // at distance 2 and 0.781 without); the addresses are this channel's plus one stride
      if (c + 1 < C) {
        if ((lane & 3) == 0) {
          const uint16_t* np = rp + (size_t)nz * plane;
          for (int z = 0; z < nz; ++z)
#pragma unroll
            for (int r = 0; r < BIN; ++r)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(np + (size_t)z * plane + (size_t)r * W));
        }
        if (HAS_ILLUM && (lane & 1) == 0) {
          const float* nip = illum + (size_t)(c + 1) * plane + (size_t)y0 * W + x0;
#pragma unroll
          for (int r = 0; r < BIN; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(nip + (size_t)r * W));
        }
      }
      if (ZT > 0) {
        uint4 v[BIN][ZT > 0 ? ZT : 1];
#pragma unroll
        for (int r = 0; r < BIN; ++r)
#pragma unroll
          for (int z = 0; z < ZT; ++z) v[r][z] = ldg128_stream(rp + (size_t)z * plane + (size_t)r * W, pol_stream);
        if (HAS_ILLUM) {
          const float* ip = illum + (size_t)c * plane + (size_t)y0 * W + x0;
#pragma unroll
          for (int r = 0; r < BIN; ++r) {
            il[r][0] = ldg128_keep(ip + (size_t)r * W, pol_keep);
            il[r][1] = ldg128_keep(ip + (size_t)r * W + 4, pol_keep);
          }
        }
#pragma unroll
        for (int r = 0; r < BIN; ++r) {
          m[r] = v[r][0];
#pragma unroll
          for (int z = 1; z < ZT; ++z) m[r] = vmax_u16x8(m[r], v[r][z]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < BIN; ++r) m[r] = ldg128_stream(rp + (size_t)r * W, pol_stream);
        if (HAS_ILLUM) {
          const float* ip = illum + (size_t)c * plane + (size_t)y0 * W + x0;
#pragma unroll
          for (int r = 0; r < BIN; ++r) {
            il[r][0] = ldg128_keep(ip + (size_t)r * W, pol_keep);
            il[r][1] = ldg128_keep(ip + (size_t)r * W + 4, pol_keep);
          }
        }
        for (int z = 1; z < nz; ++z) {
#pragma unroll
          for (int r = 0; r < BIN; ++r)
            m[r] = vmax_u16x8(m[r], ldg128_stream(rp + (size_t)z * plane + (size_t)r * W, pol_stream));
        }
      }
      if (maxproj != nullptr) {
        uint16_t* mp = maxproj + fc * plane + (size_t)y0 * W + x0;
#pragma unroll
        for (int r = 0; r < BIN; ++r) stg128_stream(mp + (size_t)r * W, m[r], pol_stream);
      }
    }

    float fv[BIN][OA_PX];
    unsigned iv[BIN][OA_PX];
    constexpr int NB = OA_PX / BIN;
#pragma unroll
    for (int r = 0; r < BIN; ++r) unpack_u16x8(m[r], iv[r]);
    if (HAS_ILLUM) {
      float bs[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) bs[j] = 0.f;
#pragma unroll
      for (int r = 0; r < BIN; ++r) {
        const float d[8] = {__uint_as_float(il[r][0].x), __uint_as_float(il[r][0].y),
                            __uint_as_float(il[r][0].z), __uint_as_float(il[r][0].w),
                            __uint_as_float(il[r][1].x), __uint_as_float(il[r][1].y),
                            __uint_as_float(il[r][1].z), __uint_as_float(il[r][1].w)};
#pragma unroll
        for (int i = 0; i < OA_PX; ++i) {
          fv[r][i] = fast_div((float)iv[r][i], d[i]);
          bs[i / BIN] += fv[r][i];
        }
      }
      if (active && binned != nullptr) {
        float* bp = reinterpret_cast<float*>(binned) + fc * (plane / (BIN * BIN)) + (size_t)rb * (W / BIN) +
                    (size_t)g * NB;
        if (NB == 8) {
          stg128_stream(bp, make_uint4(__float_as_uint(bs[0]), __float_as_uint(bs[1 % NB]),
                                       __float_as_uint(bs[2 % NB]), __float_as_uint(bs[3 % NB])), pol_stream);
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
      unsigned bs[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) bs[j] = 0u;
#pragma unroll
      for (int r = 0; r < BIN; ++r)
#pragma unroll
        for (int i = 0; i < OA_PX; ++i) {
          fv[r][i] = 0.f;
          bs[i / BIN] += iv[r][i];
        }
      if (active && binned != nullptr) {
        unsigned* bp = reinterpret_cast<unsigned*>(binned) + fc * (plane / (BIN * BIN)) + (size_t)rb * (W / BIN) +
                       (size_t)g * NB;
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
    if (warp_fg) oa_channel<BIN, HAS_ILLUM>(L, c, fv, iv, sh, rec_f, C);
  }
  if (overflow) atomicOr(flags + f, 1);
  oa_finish<HAS_ILLUM>(sh, rec_f, C);
}

template <int BIN, bool HAS_ILLUM>
static void launch_fused(int Z, int grid, cudaStream_t st, const uint16_t* raw, const float* illum,
                         const int32_t* labels, uint16_t* maxproj, void* binned, unsigned long long* rec,
                         int* flags, int Nmax, int F, int C, int H, int W, int tiles_x) {
  // Occupancy is what hides the latency here (measured on B200, profiles/README.md): 64
  // registers per thread (32 warps per SM, no spills) beats 80 and 118 registers and a
  // software-pipelined 240-register variant.  BIN = 4 holds a 4 x 8 window per lane and needs
  // the larger register budget.
  constexpr int MINB = BIN == 4 ? 4 : 8;
#define IPS_FF_CASE(ZT)                                                           \
  field_fused_kernel<BIN, ZT, HAS_ILLUM, MINB><<<grid, OA_THREADS, 0, st>>>(      \
      raw, illum, labels, maxproj, binned, rec, flags, Nmax, F, C, Z, H, W, tiles_x)
  switch (Z) {
    case 3: IPS_FF_CASE(3); break;
    case 5: IPS_FF_CASE(5); break;
    default: IPS_FF_CASE(0); break;
  }
#undef IPS_FF_CASE
}

}  // namespace ips

using namespace ips;

// defined in object_stats.cu
namespace ips {
int k3_launch_init(unsigned long long* rec, int* flags, int F, int C, int Nmax, cudaStream_t st);
int k3_launch_compact(const unsigned long long* rec, const int* flags, int32_t* n_objects, int32_t* ints,
                      float* flts, int Nmax, int F, int C, float intensity_scale, bool has_illum, cudaStream_t st);
size_t k3_records_bytes_pub(int F, int C, int Nmax);
}  // namespace ips

extern "C" size_t ips_field_fused_workspace_bytes(int F, int C, int H, int W, int bin, int Nmax) {
  (void)H; (void)W; (void)bin;
  return ips_object_stats_workspace_bytes(F, C, Nmax);
}

namespace ips {
int field_fused2_try(const uint16_t* raw, const float* illum, int illum_is_rcp, const void* labels, int label_bytes,
                     uint16_t* maxproj, void* binned, int bin, float intensity_scale, int32_t* n_objects,
                     int32_t* ints, float* flts, int Nmax, void* ws, int F, int C, int Z, int H, int W,
                     cudaStream_t st);
}

static bool force_first_generation() {
  static const bool v = [] {
    const char* e = getenv("IPS_FUSED_V1");          // A/B measurements only
    return e != nullptr && e[0] == '1';
  }();
  return v;
}

static int field_fused_impl(const char* who, const uint16_t* raw, const float* illum, int illum_is_rcp,
                            const void* labels, int label_bytes, uint16_t* maxproj, void* binned, int bin,
                            float intensity_scale, int32_t* n_objects, int32_t* ints, float* flts, int Nmax, void* ws,
                            size_t ws_bytes, int F, int C, int Z, int H, int W, ips_stream_t stream) {
  if (!raw || !labels || !n_objects || !ints || !flts) IPS_FAIL(IPS_ERR_BAD_ARG, "%s: NULL pointer argument", who);
  if (F < 0 || Z <= 0 || H <= 0 || W <= 0 || Nmax <= 0)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "%s: bad shape F=%d Z=%d H=%d W=%d Nmax=%d", who, F, Z, H, W, Nmax);
  if (C < 1 || C > OA_CMAX) IPS_FAIL(IPS_ERR_BAD_SHAPE, "%s: C must be in 1..%d (got %d)", who, OA_CMAX, C);
  if (bin != 1 && bin != 2 && bin != 4) IPS_FAIL(IPS_ERR_BAD_ARG, "%s: bin must be 1, 2 or 4 (got %d)", who, bin);
  if (H % bin || W % bin) IPS_FAIL(IPS_ERR_BAD_SHAPE, "%s: %dx%d not divisible by bin %d", who, H, W, bin);
  if (label_bytes != 2 && label_bytes != 4) IPS_FAIL(IPS_ERR_BAD_DTYPE, "%s: labels must be uint16 (2) or int32 (4)", who);
  if (illum_is_rcp && illum == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "%s: reciprocal flag without a function", who);
  if (F == 0) return IPS_OK;
  const size_t need = ips_object_stats_workspace_bytes(F, C, Nmax);
  if (ws == nullptr || ws_bytes < need) IPS_FAIL(IPS_ERR_NOMEM, "%s: needs %zu workspace bytes (got %zu)", who, need, ws_bytes);
  if (!aligned16(ws)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "%s: workspace not 16-byte aligned", who);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool general_ok = label_bytes == 4 && !illum_is_rcp;
  if (!(general_ok && force_first_generation())) {
    const int did = field_fused2_try(raw, illum, illum_is_rcp, labels, label_bytes, maxproj, binned, bin, intensity_scale,
                                     n_objects, ints, flts, Nmax, ws, F, C, Z, H, W, st);
    if (did != 0) return did < 0 ? did : IPS_OK;
  }
  if (!general_ok)
    IPS_FAIL(IPS_ERR_BAD_ALIGN, "%s: uint16 label masks / a reciprocal function need W %% 8 == 0, 16-byte aligned "
             "buffers and Nmax <= 65535 (W=%d Nmax=%d)", who, W, Nmax);
  const int32_t* labels32 = reinterpret_cast<const int32_t*>(labels);
  const bool vec = (W % 8 == 0) && aligned16(raw) && aligned16(illum) && aligned16(labels) && aligned16(maxproj) &&
                   (binned == nullptr || (reinterpret_cast<uintptr_t>(binned) & (bin == 4 ? 7u : 15u)) == 0);
  if (!vec) {
    // shapes the 128-bit kernels do not take: same results from the two general kernels
    uint16_t* mp = maxproj;
    if (mp == nullptr) IPS_FAIL(IPS_ERR_BAD_ALIGN, "%s: W %% 8 != 0 or unaligned buffers need a maxproj output buffer", who);
    int rc = ips_preprocess_fused(raw, illum, mp, nullptr, binned, bin, nullptr, nullptr, 0, F, C, Z, H, W, stream);
    if (rc != IPS_OK) return rc;
    return ips_object_stats(labels32, mp, illum, intensity_scale, n_objects, ints, flts, Nmax, ws, ws_bytes, F, C, H, W, stream);
  }
  unsigned long long* rec = reinterpret_cast<unsigned long long*>(ws);
  int* flags = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + k3_records_bytes_pub(F, C, Nmax));
  int rc = k3_launch_init(rec, flags, F, C, Nmax, st);
  if (rc != IPS_OK) return rc;
  const int tiles_x = (W + 32 * OA_PX - 1) / (32 * OA_PX);
  const int tiles_y = (H / bin + OA_WARPS - 1) / OA_WARPS;
  const long blocks_l = (long)tiles_x * tiles_y * F;
  if (blocks_l > 0x7fffffffL) IPS_FAIL(IPS_ERR_BAD_SHAPE, "%s: batch too large for one launch", who);
  const int grid = (int)blocks_l;
  const bool has_illum = illum != nullptr;
#define IPS_FF_BIN(B)                                                                                         \
  do {                                                                                                        \
    if (has_illum) launch_fused<B, true>(Z, grid, st, raw, illum, labels32, maxproj, binned, rec, flags, Nmax, F, C, H, W, tiles_x); \
    else launch_fused<B, false>(Z, grid, st, raw, illum, labels32, maxproj, binned, rec, flags, Nmax, F, C, H, W, tiles_x);          \
  } while (0)
  if (bin == 1) IPS_FF_BIN(1);
  else if (bin == 2) IPS_FF_BIN(2);
  else IPS_FF_BIN(4);
#undef IPS_FF_BIN
  IPS_LAUNCH_OK("field_fused_kernel");
  return k3_launch_compact(rec, flags, n_objects, ints, flts, Nmax, F, C, intensity_scale, has_illum, st);
}

extern "C" int ips_field_fused(const uint16_t* raw, const float* illum, const int32_t* labels,
                               uint16_t* maxproj, void* binned, int bin, float intensity_scale,
                               int32_t* n_objects, int32_t* ints, float* flts, int Nmax, void* ws,
                               size_t ws_bytes, int F, int C, int Z, int H, int W, ips_stream_t stream) {
  return field_fused_impl("ips_field_fused", raw, illum, 0, labels, 4, maxproj, binned, bin, intensity_scale, n_objects,
                          ints, flts, Nmax, ws, ws_bytes, F, C, Z, H, W, stream);
}

extern "C" int ips_field_fused_ex(const uint16_t* raw, const float* illum, int illum_is_reciprocal, const void* labels,
                                  int label_bytes, uint16_t* maxproj, void* binned, int bin, float intensity_scale,
                                  int32_t* n_objects, int32_t* ints, float* flts, int Nmax, void* ws, size_t ws_bytes,
                                  int F, int C, int Z, int H, int W, ips_stream_t stream) {
  return field_fused_impl("ips_field_fused_ex", raw, illum, illum_is_reciprocal, labels, label_bytes, maxproj, binned, bin,
                          intensity_scale, n_objects, ints, flts, Nmax, ws, ws_bytes, F, C, Z, H, W, stream);
}
