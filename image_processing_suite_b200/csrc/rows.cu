// Dense per-object feature rows for the all-gather and the well aggregation.
//
// The object kernels emit padded per-field blocks (ints [F][Nmax][6], flts [F][Nmax][2+5C],
// n_objects [F]).  ips_pack_rows compacts them into the row table that the reference's
// consumers read from Nuclei.csv / Cells.csv (Normalize_CP_ami.py:57-64): one float32 row per
// object, [well, field, label, area, y0, x0, y1, x1, cy, cx, C x (sum, mean, std, min, max)],
// rows of a field contiguous, fields in input order.  Integers are exact in float32 (< 2^24).
// ips_rows_well_ids turns column 0 of a gathered [world][cap][D] table into the int32 well id
// per row (-1 for the padding behind each rank's count) that ips_well_mean takes.
#include "ips_common.cuh"

namespace ips {

// exclusive scan of max(n_objects, 0) over F fields, single block
__global__ void __launch_bounds__(1024)
rows_offsets_kernel(const int32_t* __restrict__ n_objects, int64_t* __restrict__ offsets, int64_t* __restrict__ total,
                    uint32_t* __restrict__ header, int header_words, int F) {
  __shared__ long long warp_sum[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int start = 0; start < F; start += blockDim.x) {
    const int i = start + threadIdx.x;
    long long v = (i < F && n_objects[i] > 0) ? n_objects[i] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    long long wbase = 0, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      if (w < warp) wbase += warp_sum[w];
      tot += warp_sum[w];
    }
    const long long carry = carry_s;
    if (i < F) offsets[i] = carry + wbase + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total != nullptr) *total = carry_s;
  // block header (ips_pack_rows_block): row count as two 32-bit words, the rest of the row zero
  if (header != nullptr)
    for (int k = threadIdx.x; k < header_words; k += blockDim.x)
      header[k] = k == 0 ? (uint32_t)(carry_s & 0xffffffffll) : k == 1 ? (uint32_t)(carry_s >> 32) : 0u;
}

__global__ void __launch_bounds__(256)
rows_pack_kernel(const int32_t* __restrict__ ints, const float* __restrict__ flts,
                 const int32_t* __restrict__ n_objects, const int32_t* __restrict__ field_well,
                 const int64_t* __restrict__ offsets, float* __restrict__ rows, int Nmax, int nf, int field_base) {
  const int f = blockIdx.y;
  const int n = n_objects[f];
  const int D = 8 + nf;
  const float well = (float)field_well[f];
  const float field = (float)(field_base + f);
  float* dst = rows + (size_t)offsets[f] * D;
  const int32_t* si = ints + (size_t)f * Nmax * 6;
  const float* sf = flts + (size_t)f * Nmax * nf;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)n * D;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / D), c = (int)(e - (long long)r * D);
    float v;
    if (c == 0) v = well;
    else if (c == 1) v = field;
    else if (c < 8) v = (float)si[(size_t)r * 6 + (c - 2)];
    else v = sf[(size_t)r * nf + (c - 8)];
    dst[e] = v;
  }
}

__global__ void rows_well_ids_kernel(const float* __restrict__ rows, const int64_t* __restrict__ counts,
                                     int32_t* __restrict__ well, long long cap, int D, long long n_total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  const long long r = i / cap, k = i - r * cap;
  well[i] = k < counts[r] ? (int)rows[(size_t)i * D] : -1;
}

}  // namespace ips

using namespace ips;

extern "C" size_t ips_pack_rows_workspace_bytes(int F) {
  return F > 0 ? round_up((size_t)F * sizeof(int64_t), 256) : 0;
}

extern "C" int ips_pack_rows(const int32_t* ints, const float* flts, const int32_t* n_objects,
                             const int32_t* field_well, int field_base, float* rows_out, int64_t* total_out,
                             int Nmax, int C, int F, void* ws, size_t ws_bytes, ips_stream_t stream) {
  if (!ints || !flts || !n_objects || !field_well || !rows_out || !total_out)
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_pack_rows: NULL pointer argument");
  if (F <= 0 || F > 65535 || Nmax <= 0 || C < 1 || C > 8)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_pack_rows: bad shape F=%d Nmax=%d C=%d", F, Nmax, C);
  const size_t need = ips_pack_rows_workspace_bytes(F);
  if (ws == nullptr || ws_bytes < need) IPS_FAIL(IPS_ERR_NOMEM, "ips_pack_rows: needs %zu workspace bytes (got %zu)", need, ws_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t* offsets = reinterpret_cast<int64_t*>(ws);
  rows_offsets_kernel<<<1, 1024, 0, st>>>(n_objects, offsets, total_out, nullptr, 0, F);
  IPS_LAUNCH_OK("rows_offsets_kernel");
  const int nf = 2 + 5 * C;
  const long long per_field = (long long)Nmax * (8 + nf);
  const int bx = (int)((per_field + 256 * 8 - 1) / (256 * 8));
  rows_pack_kernel<<<dim3(bx > 0 ? bx : 1, F), 256, 0, st>>>(ints, flts, n_objects, field_well, offsets, rows_out,
                                                            Nmax, nf, field_base);
  IPS_LAUNCH_OK("rows_pack_kernel");
  return IPS_OK;
}

// Same rows behind a header row: block_out [block_rows][D]; row 0 = header (row count as two
// 32-bit words, rest zero), rows 1 .. count = the objects.  This is the unit ips_allgather_blocks
// moves: the count travels inside the block, so neither side needs a host round trip.
extern "C" int ips_pack_rows_block(const int32_t* ints, const float* flts, const int32_t* n_objects,
                                   const int32_t* field_well, int field_base, float* block_out, int64_t block_rows,
                                   int Nmax, int C, int F, void* ws, size_t ws_bytes, ips_stream_t stream) {
  if (!ints || !flts || !n_objects || !field_well || !block_out)
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_pack_rows_block: NULL pointer argument");
  if (F <= 0 || F > 65535 || Nmax <= 0 || C < 1 || C > 8)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_pack_rows_block: bad shape F=%d Nmax=%d C=%d", F, Nmax, C);
  if (block_rows - 1 < (int64_t)F * Nmax)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_pack_rows_block: block of %lld rows cannot hold a header and %d x %d objects",
             (long long)block_rows, F, Nmax);
  const size_t need = ips_pack_rows_workspace_bytes(F);
  if (ws == nullptr || ws_bytes < need) IPS_FAIL(IPS_ERR_NOMEM, "ips_pack_rows_block: needs %zu workspace bytes (got %zu)", need, ws_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t* offsets = reinterpret_cast<int64_t*>(ws);
  const int nf = 2 + 5 * C, D = 8 + nf;
  rows_offsets_kernel<<<1, 1024, 0, st>>>(n_objects, offsets, nullptr, reinterpret_cast<uint32_t*>(block_out), D, F);
  IPS_LAUNCH_OK("rows_offsets_kernel");
  const long long per_field = (long long)Nmax * D;
  const int bx = (int)((per_field + 256 * 8 - 1) / (256 * 8));
  rows_pack_kernel<<<dim3(bx > 0 ? bx : 1, F), 256, 0, st>>>(ints, flts, n_objects, field_well, offsets, block_out + D,
                                                            Nmax, nf, field_base);
  IPS_LAUNCH_OK("rows_pack_kernel");
  return IPS_OK;
}

// counts_out[b] (device int64) = header count of block b of a [n_blocks][block_rows][D] table.
namespace ips {
__global__ void block_counts_kernel(const float* __restrict__ table, int64_t* __restrict__ counts, long long n_blocks,
                                    long long block_rows, int D) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const uint32_t* h = reinterpret_cast<const uint32_t*>(table + (size_t)b * block_rows * D);
  counts[b] = (long long)h[0] | ((long long)h[1] << 32);
}
}  // namespace ips

extern "C" int ips_block_counts(const float* table, int64_t* counts_out, int64_t n_blocks, int64_t block_rows, int D,
                                ips_stream_t stream) {
  if (!table || !counts_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_block_counts: NULL pointer argument");
  if (n_blocks <= 0 || block_rows < 1 || D < 2) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_block_counts: bad shape");
  block_counts_kernel<<<(unsigned)((n_blocks + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      table, counts_out, n_blocks, block_rows, D);
  IPS_LAUNCH_OK("block_counts_kernel");
  return IPS_OK;
}

extern "C" int ips_rows_well_ids(const float* rows, const int64_t* counts_dev, int32_t* well_out,
                                 int64_t cap_per_rank, int world, int D, ips_stream_t stream) {
  if (!rows || !counts_dev || !well_out) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_rows_well_ids: NULL pointer argument");
  if (cap_per_rank <= 0 || world <= 0 || D <= 0) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_rows_well_ids: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = (long long)cap_per_rank * world;
  rows_well_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rows, counts_dev, well_out, cap_per_rank, D, n);
  IPS_LAUNCH_OK("rows_well_ids_kernel");
  return IPS_OK;
}
