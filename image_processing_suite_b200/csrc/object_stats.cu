// K3 -- per-object statistics over a label mask: a label-keyed segmented reduction.
//
// Replaces the CellProfiler MeasureObjectSizeShape / MeasureObjectIntensity subprocess
// (Feature_extraction_opt.py:166-167) and the regionprops conventions of
// Cellpose_GPU_s3fs.py:149-170: area, half-open bbox, centroid, and per channel
// sum (integrated intensity), mean, population std, min, max.
//
// Three launches per batch of fields:
//   1. init      -- reset the per-(field, label) accumulator records.
//   2. scan      -- one streaming pass over labels, max projection and illumination function:
//                   every lane reads a 2 x 8 pixel window with 128-bit loads; the label-keyed
//                   reduction runs lane -> warp (segmented shuffle tree) -> CTA (record list
//                   in shared memory) -> global accumulators (atomics, once per object and
//                   256 x 8 tile); see object_accum.cuh.
//   3. compact   -- prefix-sum over "area > 0" and emit dense rows in ascending label order,
//                   converting sums to mean / std / centroid (256 labels per block).
// Blocks are numbered field-fastest so that the blocks reading the same piece of the
// plate-constant illumination function run together and share it through L2.
#include "object_accum.cuh"

namespace ips {

constexpr int K3_ROWS = 2;

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

// One channel of a lane's window as it comes out of memory.
struct K3ChanRegs {
  uint4 px[K3_ROWS];
  uint4 il[K3_ROWS][2];
};

template <bool HAS_ILLUM, bool VEC>
__device__ __forceinline__ void k3_load_chan(K3ChanRegs& R, const uint16_t* __restrict__ mp,
                                             const float* __restrict__ illum, size_t chan_off, int y0, int x0,
                                             int H, int W, unsigned row_need, uint64_t pol_stream,
                                             uint64_t pol_keep) {
#pragma unroll
  for (int r = 0; r < K3_ROWS; ++r) {
    R.px[r] = make_uint4(0u, 0u, 0u, 0u);
    R.il[r][0] = R.il[r][1] = make_uint4(0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u);
    const int y = y0 + r;
    if (((row_need >> r) & 1u) && y < H) {
      const size_t off = chan_off + (size_t)y * W + x0;
      if (VEC) {
        R.px[r] = ldg128_stream(mp + off, pol_stream);
        if (HAS_ILLUM) {
          R.il[r][0] = ldg128_keep(illum + off, pol_keep);
          R.il[r][1] = ldg128_keep(illum + off + 4, pol_keep);
        }
      } else {
        unsigned v[OA_PX];
        unsigned d[OA_PX];
#pragma unroll
        for (int i = 0; i < OA_PX; ++i) {
          const bool in = x0 + i < W;
          v[i] = in ? (unsigned)mp[off + i] : 0u;
          d[i] = (HAS_ILLUM && in) ? __float_as_uint(illum[off + i]) : 0x3f800000u;
        }
        R.px[r] = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
        R.il[r][0] = make_uint4(d[0], d[1], d[2], d[3]);
        R.il[r][1] = make_uint4(d[4], d[5], d[6], d[7]);
      }
    }
  }
}

template <bool HAS_ILLUM, bool VEC>
__global__ void __launch_bounds__(OA_THREADS, 8)
object_stats_scan_kernel(const int32_t* __restrict__ labels, const uint16_t* __restrict__ maxproj,
                         const float* __restrict__ illum, unsigned long long* __restrict__ rec,
                         int* __restrict__ flags, int Nmax, int F, int C, int H, int W, int tiles_x) {
  __shared__ OaShared sh;
  oa_init_shared(sh);
  const int bid = blockIdx.x;
  const int f = bid % F;
  const int t = bid / F;
  const int tile_x = t % tiles_x, tile_y = t / tiles_x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y0 = (tile_y * OA_WARPS + warp) * K3_ROWS;
  const int x0 = (tile_x * 32 + lane) * OA_PX;
  const size_t plane = (size_t)H * W;
  const int32_t* lp = labels + (size_t)f * plane;
  const uint16_t* mp = maxproj + (size_t)f * C * plane;
  unsigned long long* rec_f = rec + (size_t)f * Nmax * k3_record_words(C);
  const uint64_t pol_stream = policy_evict_first();
  const uint64_t pol_keep = policy_evict_last();
  const bool col_ok = x0 < W;

  int lab[K3_ROWS][OA_PX];
#pragma unroll
  for (int r = 0; r < K3_ROWS; ++r) {
    const int y = y0 + r;
    if (col_ok && y < H) {
      const size_t off = (size_t)y * W + x0;
      if (VEC) {
        const uint4 l0 = ldg128_stream(lp + off, pol_stream);
        const uint4 l1 = ldg128_stream(lp + off + 4, pol_stream);
        lab[r][0] = (int)l0.x; lab[r][1] = (int)l0.y; lab[r][2] = (int)l0.z; lab[r][3] = (int)l0.w;
        lab[r][4] = (int)l1.x; lab[r][5] = (int)l1.y; lab[r][6] = (int)l1.z; lab[r][7] = (int)l1.w;
      } else {
#pragma unroll
        for (int i = 0; i < OA_PX; ++i) lab[r][i] = (x0 + i < W) ? lp[off + i] : 0;
      }
    } else {
#pragma unroll
      for (int i = 0; i < OA_PX; ++i) lab[r][i] = 0;
    }
  }
  K3ChanRegs cur;

  bool overflow = false;
  OaLane<K3_ROWS> L;
  oa_begin<K3_ROWS>(L, lab, lp + (size_t)y0 * W + x0, W, Nmax, y0, x0, sh, rec_f, C, overflow);
  const unsigned fg = L.m1 | L.m2 | L.m3;
  unsigned row_need = 0u;
#pragma unroll
  for (int r = 0; r < K3_ROWS; ++r) row_need |= ((fg >> (r * OA_PX)) & 0xffu) ? (1u << r) : 0u;

  auto consume = [&](const K3ChanRegs& R, int c) {
    float fv[K3_ROWS][OA_PX];
    unsigned iv[K3_ROWS][OA_PX];
#pragma unroll
    for (int r = 0; r < K3_ROWS; ++r) {
      unpack_u16x8(R.px[r], iv[r]);
      if (HAS_ILLUM) {
        const float d[8] = {__uint_as_float(R.il[r][0].x), __uint_as_float(R.il[r][0].y),
                            __uint_as_float(R.il[r][0].z), __uint_as_float(R.il[r][0].w),
                            __uint_as_float(R.il[r][1].x), __uint_as_float(R.il[r][1].y),
                            __uint_as_float(R.il[r][1].z), __uint_as_float(R.il[r][1].w)};
#pragma unroll
        for (int i = 0; i < OA_PX; ++i) fv[r][i] = fast_div((float)iv[r][i], d[i]);
      } else {
#pragma unroll
        for (int i = 0; i < OA_PX; ++i) fv[r][i] = 0.f;
      }
    }
    oa_channel<K3_ROWS, HAS_ILLUM>(L, c, fv, iv, sh, rec_f, C);
  };
  // Occupancy (8 four-warp CTAs per SM at 64 registers) hides the load latency better than a second
  // register set would (measured, profiles/README.md).  All-background rows are not loaded.
  if (__any_sync(OA_FULL, fg != 0u)) {
    for (int c = 0; c < C; ++c) {
      // next channel's rows towards L2 while this one is folded (as in field_fused.cu)
      if (VEC && c + 1 < C && row_need != 0u) {
#pragma unroll
        for (int r = 0; r < K3_ROWS; ++r) {
          if (!((row_need >> r) & 1u) || y0 + r >= H) continue;
          const size_t off = (size_t)(c + 1) * plane + (size_t)(y0 + r) * W + x0;
          if ((lane & 3) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(mp + off));
          if (HAS_ILLUM && (lane & 1) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(illum + off));
        }
      }
      k3_load_chan<HAS_ILLUM, VEC>(cur, mp, illum, (size_t)c * plane, y0, x0, H, W, row_need, pol_stream, pol_keep);
      consume(cur, c);
    }
  }
  if (overflow) atomicOr(flags + f, 1);
  oa_finish<HAS_ILLUM>(sh, rec_f, C);
}

// Dense rows in ascending label order.  Block (chunk, field) owns K3_CHUNK consecutive labels:
// it counts the present labels before its chunk (that is its first row), ranks its own present
// labels with a ballot prefix and converts sums to mean / std / centroid.  The last chunk
// writes the field's object count.
constexpr int K3_CHUNK = 256;

template <bool HAS_ILLUM>
__global__ void __launch_bounds__(K3_CHUNK)
object_stats_compact_kernel(const unsigned long long* __restrict__ rec, const int* __restrict__ flags,
                            int32_t* __restrict__ n_objects, int32_t* __restrict__ ints,
                            float* __restrict__ flts, int Nmax, int C, float intensity_scale) {
  const int f = blockIdx.y;
  const int chunk = blockIdx.x;
  const int words = k3_record_words(C);
  const int nf = 2 + 5 * C;
  const unsigned long long* rec_f = rec + (size_t)f * Nmax * words;
  int32_t* ints_f = ints + (size_t)f * Nmax * 6;
  float* flts_f = flts + (size_t)f * Nmax * nf;
  __shared__ int warp_cnt[K3_CHUNK / 32];
  __shared__ int warp_tot[K3_CHUNK / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // rows taken by the labels of earlier chunks
  int before = 0;
  for (int li = threadIdx.x; li < chunk * K3_CHUNK; li += K3_CHUNK)
    before += reinterpret_cast<const unsigned*>(rec_f + (size_t)li * words)[4] > 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 0) warp_cnt[warp] = before;

  const int li = chunk * K3_CHUNK + threadIdx.x;  // label - 1
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
  int base = 0, wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < K3_CHUNK / 32; ++w) {
    base += warp_cnt[w];
    if (w < warp) wbase += warp_tot[w];
    total += warp_tot[w];
  }
  const double sc = (double)intensity_scale;
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
  if (chunk == (int)gridDim.x - 1 && threadIdx.x == 0) n_objects[f] = flags[f] ? -1 : base + total;
}

static size_t k3_records_bytes(int F, int C, int Nmax) {
  return round_up((size_t)F * Nmax * k3_record_words(C) * 8, 256);
}
size_t k3_records_bytes_pub(int F, int C, int Nmax) { return k3_records_bytes(F, C, Nmax); }

// shared with field_fused.cu
int k3_launch_init(unsigned long long* rec, int* flags, int F, int C, int Nmax, cudaStream_t st) {
  const int words = k3_record_words(C);
  const size_t n_records = (size_t)F * Nmax;
  const size_t n = n_records * words;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  object_stats_init_kernel<<<blocks, 256, 0, st>>>(rec, flags, n_records, words, C, F);
  IPS_LAUNCH_OK("object_stats_init_kernel");
  return IPS_OK;
}

int k3_launch_compact(const unsigned long long* rec, const int* flags, int32_t* n_objects, int32_t* ints,
                      float* flts, int Nmax, int F, int C, float intensity_scale, bool has_illum, cudaStream_t st) {
  if (F > 65535) IPS_FAIL(IPS_ERR_BAD_SHAPE, "object statistics: at most 65535 fields per call (got %d)", F);
  if (has_illum)
    object_stats_compact_kernel<true><<<dim3((Nmax + K3_CHUNK - 1) / K3_CHUNK, F), K3_CHUNK, 0, st>>>(rec, flags, n_objects, ints, flts, Nmax, C, intensity_scale);
  else
    object_stats_compact_kernel<false><<<dim3((Nmax + K3_CHUNK - 1) / K3_CHUNK, F), K3_CHUNK, 0, st>>>(rec, flags, n_objects, ints, flts, Nmax, C, intensity_scale);
  IPS_LAUNCH_OK("object_stats_compact_kernel");
  return IPS_OK;
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
  {
    const int rc = k3_launch_init(rec, flags, F, C, Nmax, st);
    if (rc != IPS_OK) return rc;
  }
  const bool vec = (W % 8 == 0) && aligned16(labels) && aligned16(maxproj) && aligned16(illum);
  const int tiles_x = (W + 32 * OA_PX - 1) / (32 * OA_PX);
  const int tiles_y = (H + OA_WARPS * K3_ROWS - 1) / (OA_WARPS * K3_ROWS);
  const long blocks_l = (long)tiles_x * tiles_y * F;
  if (blocks_l > 0x7fffffffL) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_object_stats: batch too large for one launch");
  const int blocks = (int)blocks_l;
  const bool has_illum = illum != nullptr;
#define IPS_K3_LAUNCH(HI, VE)                                                                       \
  object_stats_scan_kernel<HI, VE><<<blocks, OA_THREADS, 0, st>>>(labels, maxproj, illum, rec, flags, \
                                                                  Nmax, F, C, H, W, tiles_x)
  if (has_illum) { if (vec) IPS_K3_LAUNCH(true, true); else IPS_K3_LAUNCH(true, false); }
  else { if (vec) IPS_K3_LAUNCH(false, true); else IPS_K3_LAUNCH(false, false); }
#undef IPS_K3_LAUNCH
  IPS_LAUNCH_OK("object_stats_scan_kernel");
  return k3_launch_compact(rec, flags, n_objects, ints, flts, Nmax, F, C, intensity_scale, has_illum, st);
}
