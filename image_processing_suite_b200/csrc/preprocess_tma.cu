// K1, TMA-staged variant: the same fused z-max -> illumination divide -> b x b sum-bin pass as
// preprocess.cu, with the row tiles brought into shared memory by the TMA unit
// (cp.async.bulk, mbarrier complete_tx) instead of per-lane 128-bit loads.
//
// Persistent CTAs, one per SM: warp 0 is the producer -- for every work item (one row block of
// one (field, channel) plane) it issues Z x BIN bulk copies of raw rows and BIN of the
// illumination function into a multi-stage ring; 9 consumer warps wait on the stage's
// mbarrier, read their 16-byte words with LDS.128, do the packed-uint16 max / divide / bin and
// store the outputs straight to global memory (they are written once and never re-read), then
// release the stage.  Work items are numbered field-fastest, so the CTAs that are in flight
// together read the same rows of the plate-constant function (L2, evict_last hint).
//
// Both variants are kept: the direct-load kernel (preprocess.cu) already runs at the measured
// HBM copy rate, so this one is selected with IPS_K1_TMA=1 and exists to compare the two
// staging strategies on the same pass (numbers in profiles/README.md).
#include "ips_common.cuh"

namespace ips {

constexpr int KT_CONSUMER_WARPS = 9;
constexpr int KT_THREADS = 32 * (1 + KT_CONSUMER_WARPS);
constexpr int KT_MAX_STAGES = 4;

__device__ __forceinline__ uint32_t kt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kt_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void kt_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void kt_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(kt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void kt_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(kt_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void kt_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(kt_smem_u32(dst)), "l"(src), "r"(bytes), "r"(kt_smem_u32(bar)), "l"(pol)
      : "memory");
}

template <int BIN, bool HAS_ILLUM>
__global__ void __launch_bounds__(KT_THREADS, 1)
preprocess_tma_kernel(const uint16_t* __restrict__ raw, const float* __restrict__ illum,
                      uint16_t* __restrict__ maxproj, void* __restrict__ binned, int F, int C, int Z, int H,
                      int W, int stages, long long n_items) {
  extern __shared__ __align__(128) uint8_t kt_smem[];
  const uint32_t raw_row = (uint32_t)W * 2u, ill_row = (uint32_t)W * 4u;
  const uint32_t stage_bytes = BIN * (Z * raw_row + (HAS_ILLUM ? ill_row : 0u));
  uint64_t* full = reinterpret_cast<uint64_t*>(kt_smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + KT_MAX_STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      kt_mbar_init(&full[s], 1);
      kt_mbar_init(&empty[s], KT_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t plane = (size_t)H * W;
  const int n_rb = H / BIN;

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      const uint64_t pol_keep = policy_evict_last();
      int stage = 0;
      uint32_t phase = 0;
      for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int f = (int)(it % F);
        const long long t = it / F;
        const int rb = (int)(t % n_rb), c = (int)(t / n_rb);
        kt_mbar_wait(&empty[stage], phase ^ 1u);
        uint8_t* st = kt_smem + (size_t)stage * stage_bytes;
        kt_mbar_expect_tx(&full[stage], stage_bytes);
        const uint16_t* rp = raw + ((size_t)f * C + c) * Z * plane + (size_t)rb * BIN * W;
        for (int z = 0; z < Z; ++z)
#pragma unroll
          for (int r = 0; r < BIN; ++r)
            kt_bulk_load(st + (size_t)(z * BIN + r) * raw_row, rp + (size_t)z * plane + (size_t)r * W, raw_row,
                         &full[stage], pol_stream);
        if (HAS_ILLUM) {
          const float* ip = illum + (size_t)c * plane + (size_t)rb * BIN * W;
#pragma unroll
          for (int r = 0; r < BIN; ++r)
            kt_bulk_load(st + (size_t)Z * BIN * raw_row + (size_t)r * ill_row, ip + (size_t)r * W, ill_row,
                         &full[stage], pol_keep);
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // ---- consumers ---------------------------------------------------------------------------
  const uint64_t pol_stream = policy_evict_first();
  const int G = W >> 3;
  const int ct = threadIdx.x - 32;
  constexpr int NB = 8 / BIN;
  int stage = 0;
  uint32_t phase = 0;
  for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int f = (int)(it % F);
    const long long t = it / F;
    const int rb = (int)(t % n_rb), c = (int)(t / n_rb);
    const size_t fc = (size_t)f * C + c;
    kt_mbar_wait(&full[stage], phase);
    const uint8_t* st = kt_smem + (size_t)stage * stage_bytes;
    for (int g = ct; g < G; g += 32 * KT_CONSUMER_WARPS) {
      uint4 m[BIN];
#pragma unroll
      for (int r = 0; r < BIN; ++r) {
        m[r] = *reinterpret_cast<const uint4*>(st + (size_t)r * raw_row + (size_t)g * 16);
        for (int z = 1; z < Z; ++z)
          m[r] = vmax_u16x8(m[r], *reinterpret_cast<const uint4*>(st + (size_t)(z * BIN + r) * raw_row + (size_t)g * 16));
      }
      if (maxproj != nullptr) {
        uint16_t* mp = maxproj + fc * plane + (size_t)rb * BIN * W + (size_t)g * 8;
#pragma unroll
        for (int r = 0; r < BIN; ++r) stg128_stream(mp + (size_t)r * W, m[r], pol_stream);
      }
      if (binned != nullptr) {
        unsigned bsu[NB];
        if (HAS_ILLUM) {
          float bs[NB];
#pragma unroll
          for (int j = 0; j < NB; ++j) bs[j] = 0.f;
#pragma unroll
          for (int r = 0; r < BIN; ++r) {
            const uint8_t* ir = st + (size_t)Z * BIN * raw_row + (size_t)r * ill_row + (size_t)g * 32;
            const uint4 i0 = *reinterpret_cast<const uint4*>(ir);
            const uint4 i1 = *reinterpret_cast<const uint4*>(ir + 16);
            const float d[8] = {__uint_as_float(i0.x), __uint_as_float(i0.y), __uint_as_float(i0.z),
                                __uint_as_float(i0.w), __uint_as_float(i1.x), __uint_as_float(i1.y),
                                __uint_as_float(i1.z), __uint_as_float(i1.w)};
            float x[8];
            unpack_u16x8(m[r], x);
#pragma unroll
            for (int i = 0; i < 8; ++i) bs[i / BIN] += fast_div(x[i], d[i]);
          }
#pragma unroll
          for (int j = 0; j < NB; ++j) bsu[j] = __float_as_uint(bs[j]);
        } else {
#pragma unroll
          for (int j = 0; j < NB; ++j) bsu[j] = 0u;
#pragma unroll
          for (int r = 0; r < BIN; ++r) {
            uint32_t x[8];
            unpack_u16x8(m[r], x);
#pragma unroll
            for (int i = 0; i < 8; ++i) bsu[i / BIN] += x[i];
          }
        }
        unsigned* bp = reinterpret_cast<unsigned*>(binned) + fc * (plane / (BIN * BIN)) + (size_t)rb * (W / BIN) +
                       (size_t)g * NB;
        if (NB == 8) {
          stg128_stream(bp, make_uint4(bsu[0], bsu[1 % NB], bsu[2 % NB], bsu[3 % NB]), pol_stream);
          stg128_stream(bp + 4, make_uint4(bsu[4 % NB], bsu[5 % NB], bsu[6 % NB], bsu[7 % NB]), pol_stream);
        } else if (NB == 4) {
          stg128_stream(bp, make_uint4(bsu[0], bsu[1 % NB], bsu[2 % NB], bsu[3 % NB]), pol_stream);
        } else {
          stg64_stream(bp, make_uint2(bsu[0], bsu[1 % NB]), pol_stream);
        }
      }
    }
    __syncwarp();
    if (lane == 0) kt_mbar_arrive(&empty[stage]);   // this warp is done reading the stage
    if (++stage == stages) { stage = 0; phase ^= 1u; }
  }
}

// Returns IPS_OK when launched, 1 when the shape does not fit this variant (caller falls back
// to the direct-load kernel), or a negative error.
int preprocess_tma_try(const uint16_t* raw, const float* illum, uint16_t* maxproj, void* binned, int bin, int F,
                       int C, int Z, int H, int W, cudaStream_t st) {
  const size_t stage_bytes = (size_t)bin * ((size_t)Z * W * 2 + (illum ? (size_t)W * 4 : 0));
  if (W % 8 || stage_bytes == 0) return 1;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > KT_MAX_STAGES) stages = KT_MAX_STAGES;
  if (stages < 2) return 1;
  const size_t smem = (size_t)stages * stage_bytes + 2 * KT_MAX_STAGES * sizeof(uint64_t);
  const long long n_items = (long long)F * C * (H / bin);
  const int grid = (int)(n_items < (long long)sm_count() ? n_items : (long long)sm_count());
#define IPS_KT_LAUNCH(B, HI)                                                                                     \
  do {                                                                                                           \
    IPS_CUDA_OK(cudaFuncSetAttribute(preprocess_tma_kernel<B, HI>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                     (int)smem));                                                                \
    preprocess_tma_kernel<B, HI><<<grid, KT_THREADS, smem, st>>>(raw, illum, maxproj, binned, F, C, Z, H, W,     \
                                                                 stages, n_items);                               \
  } while (0)
  if (illum != nullptr) {
    if (bin == 1) IPS_KT_LAUNCH(1, true); else if (bin == 2) IPS_KT_LAUNCH(2, true); else IPS_KT_LAUNCH(4, true);
  } else {
    if (bin == 1) IPS_KT_LAUNCH(1, false); else if (bin == 2) IPS_KT_LAUNCH(2, false); else IPS_KT_LAUNCH(4, false);
  }
#undef IPS_KT_LAUNCH
  IPS_LAUNCH_OK("preprocess_tma_kernel");
  return IPS_OK;
}

}  // namespace ips
