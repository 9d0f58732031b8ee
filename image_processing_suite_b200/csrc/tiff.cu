// K7: TIFF strip codec on the device -- the file edge of Image_re-binning.py:17-21
// (Image.open(...) / img.save(..., format='TIFF', compression='tiff_lzw')) and of
// MaxProjection.py:39 (imageio.imread of each plane).
//
// ips_tiff_lzw_encode_u16: uint16 planes in HBM -> complete little-endian TIFF files in HBM,
// byte-identical to what Pillow 12.2 / libtiff 4.7 writes for the same pixels (strip LZW
// streams, strip size, tag set and placement), so only compressed bytes cross PCIe.
//   1. tiff_lzw_encode_kernel   one warp per strip, 32 bytes per step (one speculative dictionary
//                               walk per lane), hash table in 21.5 KB of shared memory, strips
//                               land in fixed-capacity slots of the workspace;
//   2. tiff_layout_kernel       one CTA per plane: scan of the strip sizes -> offsets, header,
//                               IFD and the two strip arrays written behind the strips;
//   3. tiff_gather_kernel       one CTA per strip: slot -> its place in the file.
// ips_tiff_lzw_decode: LZW strips (anywhere in a device buffer) -> pixels; one warp per strip,
// 32 codes parsed per step (one per lane), an 8.4 KB table of output offsets per generation and
// an 8 KB output window in shared memory, written out in 16-byte vectors (tiff_lzw_core.cuh).
// ips_tiff_fix_u16 undoes big-endian samples and horizontal differencing (Predictor = 2).
//
// Both directions are bound by dependent shared-memory / shuffle latency on one warp per strip,
// not by HBM (DRAM traffic is the data once); throughput scales with the resident strips
// (10 encoder or 13 decoder warps per SM) x planes x strips (36 per 1080^2 plane).
#include "ips_common.cuh"
#include "tiff_lzw_core.cuh"

namespace ips {

namespace lz = ips_lzw;

__global__ void __launch_bounds__(32)
tiff_lzw_encode_kernel(const uint8_t* __restrict__ planes, uint8_t* __restrict__ slots, uint32_t* __restrict__ strip_bytes,
                       size_t plane_bytes, uint32_t row_bytes, int H, int rps, int S, uint32_t cap) {
  __shared__ __align__(16) uint32_t table[lz::ENC_SLOTS];
  __shared__ uint32_t stage[lz::PE_STAGE];
  const int s = blockIdx.x, p = blockIdx.y;
  const int r0 = s * rps;
  const int rows = min(rps, H - r0);
  const uint8_t* in = planes + (size_t)p * plane_bytes + (size_t)r0 * row_bytes;
  const size_t slot = (size_t)p * S + s;
  lz::Warp w;
  const uint32_t n = lz::encode_strip(in, (uint32_t)rows * row_bytes, slots + slot * cap, cap, table, stage, w);
  if (w.lane == 0) strip_bytes[slot] = n;
}

__device__ __forceinline__ void put16le(uint8_t* p, uint32_t v) {
  p[0] = (uint8_t)v;
  p[1] = (uint8_t)(v >> 8);
}
__device__ __forceinline__ void put32le(uint8_t* p, uint32_t v) {
  p[0] = (uint8_t)v;
  p[1] = (uint8_t)(v >> 8);
  p[2] = (uint8_t)(v >> 16);
  p[3] = (uint8_t)(v >> 24);
}
__device__ __forceinline__ void put_tag(uint8_t* p, uint32_t tag, uint32_t type, uint32_t count, uint32_t value) {
  put16le(p, tag);
  put16le(p + 2, type);
  put32le(p + 4, count);
  put32le(p + 8, value);
}

// Pillow's layout: "II*\0" + IFD offset, strips from byte 8, pad to even, IFD with 9 tags
// (256 257 258 259 262 273 278 279 284), then StripByteCounts[S], then StripOffsets[S].
__global__ void __launch_bounds__(256)
tiff_layout_kernel(const uint32_t* __restrict__ strip_bytes, uint32_t* __restrict__ strip_off, uint8_t* __restrict__ files,
                   uint64_t* __restrict__ file_bytes, size_t file_cap, int H, int W, int rps, int S) {
  __shared__ uint32_t warp_sum[8];
  __shared__ uint32_t carry_s;
  __shared__ int bad_s;
  const int p = blockIdx.x;
  const uint32_t* sb = strip_bytes + (size_t)p * S;
  uint32_t* so = strip_off + (size_t)p * S;
  uint8_t* f = files + (size_t)p * file_cap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    carry_s = 8;
    bad_s = 0;
  }
  __syncthreads();
  for (int start = 0; start < S; start += 256) {
    const int i = start + threadIdx.x;
    uint32_t v = i < S ? sb[i] : 0u;
    if (v == lz::OVERFLOW) {
      bad_s = 1;
      v = 0;
    }
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
    for (int k = 0; k < 8; ++k) {
      if (k < warp) wbase += warp_sum[k];
      tot += warp_sum[k];
    }
    const uint32_t carry = carry_s;
    if (i < S) so[i] = carry + wbase + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  const uint32_t body_end = carry_s;
  const uint32_t ifd = (body_end + 1u) & ~1u;
  const uint32_t after = ifd + 2 + 12 * 9 + 4;
  const uint32_t counts_at = after, offsets_at = after + (S > 1 ? 4u * S : 0u);
  const uint64_t total = (uint64_t)offsets_at + (S > 1 ? 4ull * S : 0ull);
  if (bad_s || total > file_cap) {
    if (threadIdx.x == 0) file_bytes[p] = 0;      // reported as IPS_ERR_NOMEM by the host wrapper's caller
    return;
  }
  if (threadIdx.x == 0) {
    f[0] = 'I';
    f[1] = 'I';
    put16le(f + 2, 42);
    put32le(f + 4, ifd);
    if (ifd != body_end) f[body_end] = 0;
    uint8_t* d = f + ifd;
    put16le(d, 9);
    d += 2;
    put_tag(d + 0 * 12, 256, 3, 1, (uint32_t)W);
    put_tag(d + 1 * 12, 257, 3, 1, (uint32_t)H);
    put_tag(d + 2 * 12, 258, 3, 1, 16);
    put_tag(d + 3 * 12, 259, 3, 1, 5);
    put_tag(d + 4 * 12, 262, 3, 1, 1);
    put_tag(d + 5 * 12, 273, 4, (uint32_t)S, S > 1 ? offsets_at : 8u);
    put_tag(d + 6 * 12, 278, 3, 1, (uint32_t)rps);
    put_tag(d + 7 * 12, 279, 4, (uint32_t)S, S > 1 ? counts_at : sb[0]);
    put_tag(d + 8 * 12, 284, 3, 1, 1);
    put32le(d + 9 * 12, 0);
    file_bytes[p] = total;
  }
  if (S > 1) {
    for (int i = threadIdx.x; i < S; i += 256) {
      put32le(f + counts_at + 4 * i, sb[i]);
      put32le(f + offsets_at + 4 * i, so[i]);     // written by this thread's own earlier iteration or after a barrier
    }
  }
}

__global__ void __launch_bounds__(256)
tiff_gather_kernel(const uint8_t* __restrict__ slots, const uint32_t* __restrict__ strip_bytes,
                   const uint32_t* __restrict__ strip_off, const uint64_t* __restrict__ file_bytes,
                   uint8_t* __restrict__ files, size_t file_cap, int S, uint32_t cap) {
  const int s = blockIdx.x, p = blockIdx.y;
  if (file_bytes[p] == 0) return;
  const size_t slot = (size_t)p * S + s;
  const uint32_t n = strip_bytes[slot];
  const uint8_t* src = slots + slot * cap;
  uint8_t* dst = files + (size_t)p * file_cap + strip_off[slot];
  // head bytes to a 4-byte boundary of the destination, then words assembled from two aligned
  // source words, then the tail
  uint32_t head = (uint32_t)((4u - (reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u);
  if (head > n) head = n;
  if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
  const uint32_t words = (n - head) >> 2;
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);     // slot is 16-byte aligned
  uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + head);
  const uint32_t sh = head * 8;
  for (uint32_t i = threadIdx.x; i < words; i += 256) {
    const uint32_t lo = s32[i];
    const uint32_t hi = sh ? s32[i + 1] : 0u;                       // in the slot: cap >= n + 16
    d32[i] = sh ? __funnelshift_r(lo, hi, sh) : lo;
  }
  const uint32_t done = head + (words << 2);
  if (threadIdx.x < n - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
}

__global__ void __launch_bounds__(32)
tiff_lzw_decode_kernel(const uint8_t* __restrict__ src, const uint64_t* __restrict__ src_off,
                       const uint32_t* __restrict__ src_bytes, uint8_t* __restrict__ dst,
                       const uint64_t* __restrict__ dst_off, const uint32_t* __restrict__ dst_bytes,
                       int32_t* __restrict__ status) {
  __shared__ __align__(16) uint16_t orel[lz::PD_TAB];
  __shared__ uint32_t obase[lz::PD_TAB / 16];
  __shared__ __align__(16) uint8_t win[lz::PD_WIN];
  const int s = blockIdx.x;
  lz::Warp w;
  const int st = lz::decode_strip(src + src_off[s], src_bytes[s], dst + dst_off[s], dst_bytes[s], orel, obase, win, w);
  if (w.lane == 0) status[s] = st;
}

// rows of W uint16 samples: optional byte swap, then optional running sum modulo 2^16
__global__ void __launch_bounds__(256)
tiff_fix_u16_kernel(uint16_t* __restrict__ img, int W, int predictor, int byteswap) {
  __shared__ uint32_t part[256];
  uint16_t* row = img + (size_t)blockIdx.x * W;
  const int per = (W + 255) / 256;
  const int a = min(W, (int)threadIdx.x * per), b = min(W, a + per);
  uint32_t sum = 0;
  for (int i = a; i < b; ++i) {
    uint32_t v = row[i];
    if (byteswap) {
      v = ((v & 0xFFu) << 8) | (v >> 8);
      row[i] = (uint16_t)v;
    }
    sum += v;
  }
  if (predictor != 2) return;
  part[threadIdx.x] = sum;
  __syncthreads();
  uint32_t base = 0;
  for (int k = 0; k < (int)threadIdx.x; ++k) base += part[k];
  for (int i = a; i < b; ++i) {
    base += row[i];
    row[i] = (uint16_t)base;
  }
}

static int strips_of(int H, int rps) { return (H + rps - 1) / rps; }

}  // namespace ips

using namespace ips;

extern "C" int ips_tiff_rows_per_strip(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  const long long stride = 2ll * W;
  long long r = 65536 / stride;
  if (r < 1) r = 1;
  if (r > H) r = H;
  return (int)r;
}

extern "C" size_t ips_tiff_lzw_bound(size_t strip_bytes) { return ips_lzw::encode_bound(strip_bytes); }

extern "C" size_t ips_tiff_file_bound(int H, int W, int rows_per_strip) {
  if (H <= 0 || W <= 0 || rows_per_strip <= 0) return 0;
  const int S = strips_of(H, rows_per_strip);
  const size_t cap = ips_lzw::encode_bound((size_t)rows_per_strip * W * 2);
  return round_up(8 + (size_t)S * cap + 2 + 2 + 12 * 9 + 4 + 8 * (size_t)S, 16);
}

extern "C" size_t ips_tiff_encode_workspace_bytes(int P, int H, int W, int rows_per_strip) {
  if (P <= 0 || H <= 0 || W <= 0 || rows_per_strip <= 0) return 0;
  const size_t S = strips_of(H, rows_per_strip);
  const size_t cap = ips_lzw::encode_bound((size_t)rows_per_strip * W * 2);
  return round_up((size_t)P * S * cap, 256) + 2 * round_up((size_t)P * S * 4, 256);
}

extern "C" int ips_tiff_lzw_encode_u16(const uint16_t* planes, int P, int H, int W, int rows_per_strip, uint8_t* files,
                                       size_t file_cap, uint64_t* file_bytes, void* ws, size_t ws_bytes,
                                       ips_stream_t stream) {
  if (P < 0 || H <= 0 || W <= 0 || rows_per_strip <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_tiff_lzw_encode_u16: bad shape");
  if (P == 0) return IPS_OK;
  if (!planes || !files || !file_bytes || !ws) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_tiff_lzw_encode_u16: null pointer");
  if (!aligned16(files) || !aligned16(ws) || (file_cap & 15)) IPS_FAIL(IPS_ERR_BAD_ALIGN, "ips_tiff_lzw_encode_u16: files, ws and file_cap must be 16-byte aligned");
  if ((size_t)rows_per_strip * W * 2 >= (1ull << 28)) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_tiff_lzw_encode_u16: strips of 256 MiB or more are not supported");
  if (file_cap < ips_tiff_file_bound(H, W, rows_per_strip) || file_cap >= (1ull << 32))
    IPS_FAIL(IPS_ERR_NOMEM, "ips_tiff_lzw_encode_u16: file_cap %zu outside [ips_tiff_file_bound, 4 GiB)", file_cap);
  if (ws_bytes < ips_tiff_encode_workspace_bytes(P, H, W, rows_per_strip)) IPS_FAIL(IPS_ERR_NOMEM, "ips_tiff_lzw_encode_u16: workspace too small");
  const int S = strips_of(H, rows_per_strip);
  if (S > 65535 || P > 65535) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_tiff_lzw_encode_u16: more than 65535 strips or planes per call");
  const size_t cap = ips_lzw::encode_bound((size_t)rows_per_strip * W * 2);
  uint8_t* slots = static_cast<uint8_t*>(ws);
  uint32_t* strip_bytes = reinterpret_cast<uint32_t*>(slots + round_up((size_t)P * S * cap, 256));
  uint32_t* strip_off = strip_bytes + round_up((size_t)P * S * 4, 256) / 4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  tiff_lzw_encode_kernel<<<dim3(S, P), 32, 0, st>>>(reinterpret_cast<const uint8_t*>(planes), slots, strip_bytes,
                                                    (size_t)H * W * 2, (uint32_t)W * 2u, H, rows_per_strip, S, (uint32_t)cap);
  IPS_LAUNCH_OK("tiff_lzw_encode_kernel");
  tiff_layout_kernel<<<P, 256, 0, st>>>(strip_bytes, strip_off, files, file_bytes, file_cap, H, W, rows_per_strip, S);
  IPS_LAUNCH_OK("tiff_layout_kernel");
  tiff_gather_kernel<<<dim3(S, P), 256, 0, st>>>(slots, strip_bytes, strip_off, file_bytes, files, file_cap, S, (uint32_t)cap);
  IPS_LAUNCH_OK("tiff_gather_kernel");
  return IPS_OK;
}

extern "C" int ips_tiff_lzw_decode(const uint8_t* src, const uint64_t* src_off, const uint32_t* src_bytes, uint8_t* dst,
                                   const uint64_t* dst_off, const uint32_t* dst_bytes, int n_strips, int32_t* status,
                                   ips_stream_t stream) {
  if (n_strips < 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_tiff_lzw_decode: n_strips < 0");
  if (n_strips == 0) return IPS_OK;
  if (!src || !src_off || !src_bytes || !dst || !dst_off || !dst_bytes || !status) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_tiff_lzw_decode: null pointer");
  tiff_lzw_decode_kernel<<<n_strips, 32, 0, static_cast<cudaStream_t>(stream)>>>(src, src_off, src_bytes, dst, dst_off, dst_bytes, status);
  IPS_LAUNCH_OK("tiff_lzw_decode_kernel");
  return IPS_OK;
}

extern "C" int ips_tiff_fix_u16(uint16_t* img, int64_t rows, int W, int predictor, int byteswap, ips_stream_t stream) {
  if (rows < 0 || W <= 0 || rows > 0x7fffffffll) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_tiff_fix_u16: bad shape");
  if (predictor != 1 && predictor != 2) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_tiff_fix_u16: predictor must be 1 or 2");
  if (rows == 0 || (predictor == 1 && !byteswap)) return IPS_OK;
  if (!img) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_tiff_fix_u16: null pointer");
  tiff_fix_u16_kernel<<<(unsigned)rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(img, W, predictor, byteswap);
  IPS_LAUNCH_OK("tiff_fix_u16_kernel");
  return IPS_OK;
}
