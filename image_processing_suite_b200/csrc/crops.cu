// Centroid-centred masked cell crops with per-crop, per-channel 8-bit scaling.
//
// Replaces the per-cell Python loop of Cellpose_GPU_s3fs.py:149-182: for every object of the
// label mask (ascending label order) the integer-truncated centroid (regionprops centroid ->
// int, :157), the edge test (:159-161), the BOX x BOX crop of every channel masked by
// `mask == label` (:163-166) and `scale_to_8bit` (:34-43): 255 * (x - min) / (max - min) in
// float32, truncated; a constant crop becomes zeros.  The RGB replication of :178 is a
// broadcast of the single plane written here.
//
// Kernels:
//   crop_centroid_kernel  one warp per object: exact integer coordinate sums over the object's
//                         bounding box (from the K3 rows) -> floor(sum / area), so the truncated
//                         centroid equals int(float64 mean) exactly
//   crop_slots_kernel     per-field prefix over the kept objects (crop order = label order)
//   crop_scale_kernel     one block per (crop, channel): pass 1 min / max of the masked crop,
//                         pass 2 scale + pack four pixels per 32-bit store
#include "ips_common.cuh"

namespace ips {

__global__ void __launch_bounds__(256)
crop_centroid_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ ints,
                     const int32_t* __restrict__ n_objects, int32_t* __restrict__ cyx, int32_t* __restrict__ keep,
                     int half, int Nmax, int H, int W) {
  const int f = blockIdx.y;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n = n_objects[f];
  if (r >= Nmax) return;
  int32_t* k = keep + (size_t)f * Nmax + r;
  if (r >= n) {
    if (lane == 0) *k = 0;
    return;
  }
  const int32_t* row = ints + ((size_t)f * Nmax + r) * 6;
  const int label = row[0], y0 = row[2], x0 = row[3], y1 = row[4], x1 = row[5];
  const int32_t* lp = labels + (size_t)f * H * W;
  unsigned long long sy = 0, sx = 0, cnt = 0;
  for (int y = y0; y < y1; ++y)
    for (int x = x0 + lane; x < x1; x += 32)
      if (lp[(size_t)y * W + x] == label) { sy += (unsigned)y; sx += (unsigned)x; ++cnt; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) {
    const int yc = cnt ? (int)(sy / cnt) : 0, xc = cnt ? (int)(sx / cnt) : 0;
    cyx[((size_t)f * Nmax + r) * 2] = yc;
    cyx[((size_t)f * Nmax + r) * 2 + 1] = xc;
    *k = (cnt > 0 && yc - half >= 0 && yc + half <= H && xc - half >= 0 && xc + half <= W) ? 1 : 0;
  }
}

// one block per field: slot[r] = number of kept objects before r; kept rows (label, yc, xc)
__global__ void __launch_bounds__(1024)
crop_slots_kernel(const int32_t* __restrict__ ints, const int32_t* __restrict__ cyx, const int32_t* __restrict__ keep,
                  int32_t* __restrict__ n_kept, int32_t* __restrict__ kept, int Nmax, int max_crops) {
  const int f = blockIdx.x;
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int start = 0; start < Nmax; start += blockDim.x) {
    const int r = start + threadIdx.x;
    const bool k = r < Nmax && keep[(size_t)f * Nmax + r] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      if (w < warp) wbase += warp_tot[w];
      total += warp_tot[w];
    }
    const int base = base_s;
    if (k) {
      const int slot = base + wbase + __popc(bal & ((1u << lane) - 1u));
      if (slot < max_crops) {
        int32_t* o = kept + ((size_t)f * max_crops + slot) * 3;
        o[0] = ints[((size_t)f * Nmax + r) * 6];
        o[1] = cyx[((size_t)f * Nmax + r) * 2];
        o[2] = cyx[((size_t)f * Nmax + r) * 2 + 1];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s = base + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_kept[f] = base_s;
}

__global__ void __launch_bounds__(256)
crop_scale_kernel(const float* __restrict__ img, const int32_t* __restrict__ labels,
                  const int32_t* __restrict__ n_kept, const int32_t* __restrict__ kept,
                  uint8_t* __restrict__ crops, int box, int max_crops, int C, int H, int W) {
  const int slot = blockIdx.x, c = blockIdx.y, f = blockIdx.z;
  if (slot >= min(n_kept[f], max_crops)) return;
  const int32_t* k = kept + ((size_t)f * max_crops + slot) * 3;
  const int label = k[0], half = box >> 1;
  const int y1 = k[1] - half, x1 = k[2] - half;
  const float* ip = img + ((size_t)f * C + c) * H * W;
  const int32_t* lp = labels + (size_t)f * H * W;
  const int n = box * box;
  float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
    const int yy = idx / box, xx = idx - yy * box;
    const size_t o = (size_t)(y1 + yy) * W + (x1 + xx);
    const float v = lp[o] == label ? ip[o] : 0.f;
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  __shared__ float s_lo[8], s_hi[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  lo = s_lo[0]; hi = s_hi[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
  const float range = __fsub_rn(hi, lo);
  const bool flat = hi == lo;
  uint8_t* out = crops + (((size_t)f * max_crops + slot) * C + c) * (size_t)n;
  if ((box & 3) == 0) {
    for (int q = threadIdx.x; q < n / 4; q += blockDim.x) {
      const int idx = q * 4;
      const int yy = idx / box, xx = idx - yy * box;
      const size_t o = (size_t)(y1 + yy) * W + (x1 + xx);
      unsigned packed = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v = lp[o + j] == label ? ip[o + j] : 0.f;
        const unsigned b = flat ? 0u : (unsigned)__fdiv_rn(__fmul_rn(255.0f, __fsub_rn(v, lo)), range);
        packed |= (b & 0xffu) << (8 * j);
      }
      reinterpret_cast<unsigned*>(out)[q] = packed;
    }
  } else {
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
      const int yy = idx / box, xx = idx - yy * box;
      const size_t o = (size_t)(y1 + yy) * W + (x1 + xx);
      const float v = lp[o] == label ? ip[o] : 0.f;
      out[idx] = flat ? (uint8_t)0 : (uint8_t)__fdiv_rn(__fmul_rn(255.0f, __fsub_rn(v, lo)), range);
    }
  }
}

static size_t crops_cyx_bytes(int F, int Nmax) { return round_up((size_t)F * Nmax * 2 * sizeof(int32_t), 256); }

}  // namespace ips

using namespace ips;

extern "C" size_t ips_cell_crops_workspace_bytes(int F, int Nmax) {
  if (F <= 0 || Nmax <= 0) return 0;
  return crops_cyx_bytes(F, Nmax) + round_up((size_t)F * Nmax * sizeof(int32_t), 256);
}

extern "C" int ips_cell_crops(const float* corrected, const int32_t* labels, const int32_t* ints,
                              const int32_t* n_objects, int box, int max_crops, uint8_t* crops, int32_t* n_kept,
                              int32_t* kept, void* ws, size_t ws_bytes, int Nmax, int F, int C, int H, int W,
                              ips_stream_t stream) {
  if (!corrected || !labels || !ints || !n_objects || !crops || !n_kept || !kept)
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_cell_crops: NULL pointer argument");
  if (F <= 0 || F > 65535 || C <= 0 || C > 65535 || H <= 0 || W <= 0 || Nmax <= 0 || max_crops <= 0)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_cell_crops: bad shape F=%d C=%d H=%d W=%d Nmax=%d max_crops=%d", F, C, H, W, Nmax, max_crops);
  if (box < 2 || (box & 1) || box > H || box > W)
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_cell_crops: box must be even and fit the image (got %d for %dx%d)", box, H, W);
  const size_t need = ips_cell_crops_workspace_bytes(F, Nmax);
  if (ws == nullptr || ws_bytes < need) IPS_FAIL(IPS_ERR_NOMEM, "ips_cell_crops: needs %zu workspace bytes (got %zu)", need, ws_bytes);
  if (!aligned16(ws) || (reinterpret_cast<uintptr_t>(crops) & 3u)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_cell_crops: misaligned buffer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int32_t* cyx = reinterpret_cast<int32_t*>(ws);
  int32_t* keep = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws) + crops_cyx_bytes(F, Nmax));
  crop_centroid_kernel<<<dim3((Nmax + 7) / 8, F), 256, 0, st>>>(labels, ints, n_objects, cyx, keep, box / 2, Nmax, H, W);
  IPS_LAUNCH_OK("crop_centroid_kernel");
  crop_slots_kernel<<<F, 1024, 0, st>>>(ints, cyx, keep, n_kept, kept, Nmax, max_crops);
  IPS_LAUNCH_OK("crop_slots_kernel");
  crop_scale_kernel<<<dim3(max_crops, C, F), 256, 0, st>>>(corrected, labels, n_kept, kept, crops, box, max_crops, C, H, W);
  IPS_LAUNCH_OK("crop_scale_kernel");
  return IPS_OK;
}
