// Host-buffer pipeline: the end-to-end call a drop-in script makes for a batch of fields.
//
// One submit = H2D of the batch's raw z-stacks and label masks, the fused field pass
// (ips_field_fused: K1 + K3), D2H of max projections, binned planes and object rows.
// Three streams (copy-in, compute, copy-out) and `depth` device slots chained by events,
// so the copy-in of batch i+1 and the copy-out of batch i-1 overlap the kernels of batch
// i (PCIe is full duplex).  The illumination function is uploaded once at creation: it is
// plate-constant (Illumination_QC_mult.py:182-199 preloads it the same way).
#include <mutex>
#include <new>
#include <vector>

#include "ips_common.cuh"

namespace ips {

// uint16 label masks (what Cellpose writes when a field has < 65536 objects) go to the fused
// kernel as they are; this kernel only widens them for shapes that kernel does not take.
__global__ void __launch_bounds__(256)
widen_labels_kernel(const uint16_t* __restrict__ in, int32_t* __restrict__ out, size_t n_words /* n / 8 */,
                    size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_words) {
    const uint4 v = reinterpret_cast<const uint4*>(in)[i];
    uint4* o = reinterpret_cast<uint4*>(out) + 2 * i;
    o[0] = make_uint4(v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16);
    o[1] = make_uint4(v.z & 0xffffu, v.z >> 16, v.w & 0xffffu, v.w >> 16);
  }
  if (i == 0)
    for (size_t k = n_words * 8; k < n; ++k) out[k] = in[k];
}

struct Slot {
  uint16_t* raw = nullptr;
  int32_t* labels = nullptr;
  uint16_t* labels16 = nullptr;
  uint16_t* maxproj = nullptr;
  void* binned = nullptr;
  int32_t* n_objects = nullptr;
  int32_t* ints = nullptr;
  float* flts = nullptr;
  void* ws = nullptr;
  cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
  int64_t ticket = -1;
};

}  // namespace ips

struct ips_pipeline {
  int Fb, C, Z, H, W, bin, Nmax, depth, label_bytes;
  float scale;
  float* illum = nullptr;      // the plate's function; its reciprocal on the fast path (ips_illum_reciprocal, once)
  bool fast = false;           // shape taken by the packed-label kernel: reciprocal function, masks as they are
  bool native16 = false;       // ... and the masks are uint16
  size_t ws_bytes = 0;
  cudaStream_t s_in = nullptr, s_compute = nullptr, s_out = nullptr;
  std::vector<ips::Slot> slots;
  int64_t next_ticket = 0;
  std::mutex mu;
};

using namespace ips;

static void pipeline_free(ips_pipeline* p) {
  if (p == nullptr) return;
  for (Slot& s : p->slots) {
    cudaFree(s.raw); cudaFree(s.labels); cudaFree(s.labels16); cudaFree(s.maxproj); cudaFree(s.binned);
    cudaFree(s.n_objects); cudaFree(s.ints); cudaFree(s.flts); cudaFree(s.ws);
    if (s.h2d_done) cudaEventDestroy(s.h2d_done);
    if (s.compute_done) cudaEventDestroy(s.compute_done);
    if (s.d2h_done) cudaEventDestroy(s.d2h_done);
  }
  cudaFree(p->illum);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_compute) cudaStreamDestroy(p->s_compute);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  delete p;
}

#define PIPE_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ::ips::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,    \
                       __LINE__);                                                            \
      pipeline_free(p);                                                                      \
      return e__ == cudaErrorMemoryAllocation ? IPS_ERR_NOMEM : IPS_ERR_CUDA;                \
    }                                                                                        \
  } while (0)

extern "C" int ips_pipeline_create(ips_pipeline_t** out, int fields_per_batch, int C, int Z, int H,
                                   int W, int bin, int Nmax, int depth, int label_bytes,
                                   const float* illum_host, float intensity_scale) {
  if (out == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_pipeline_create: out is NULL");
  *out = nullptr;
  if (fields_per_batch <= 0 || C <= 0 || C > 8 || Z <= 0 || H <= 0 || W <= 0 || Nmax <= 0)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_pipeline_create: bad shape Fb=%d C=%d Z=%d H=%d W=%d Nmax=%d",
             fields_per_batch, C, Z, H, W, Nmax);
  if (bin != 1 && bin != 2 && bin != 4) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_pipeline_create: bin must be 1, 2 or 4");
  if (H % bin || W % bin) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_pipeline_create: %dx%d not divisible by bin %d", H, W, bin);
  if (depth < 1 || depth > 16) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_pipeline_create: depth must be in 1..16");
  if (label_bytes != 2 && label_bytes != 4)
    IPS_FAIL(IPS_ERR_BAD_DTYPE, "ips_pipeline_create: labels must be uint16 (2) or int32 (4), got %d bytes", label_bytes);
  ips_pipeline* p = new (std::nothrow) ips_pipeline();
  if (p == nullptr) IPS_FAIL(IPS_ERR_NOMEM, "ips_pipeline_create: out of host memory");
  p->Fb = fields_per_batch; p->C = C; p->Z = Z; p->H = H; p->W = W; p->bin = bin; p->Nmax = Nmax;
  p->depth = depth; p->scale = intensity_scale; p->label_bytes = label_bytes;
  const size_t plane = (size_t)H * W, Fb = (size_t)fields_per_batch;
  PIPE_CUDA_OK(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
  PIPE_CUDA_OK(cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking));
  PIPE_CUDA_OK(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
  if (illum_host != nullptr) {
    PIPE_CUDA_OK(cudaMalloc(&p->illum, (size_t)C * plane * sizeof(float)));
    PIPE_CUDA_OK(cudaMemcpy(p->illum, illum_host, (size_t)C * plane * sizeof(float), cudaMemcpyHostToDevice));
  }
  // The packed-label kernel takes uint16 masks directly and the function as its reciprocal; other
  // shapes go through the general kernels (masks widened on the device).
  p->fast = W % 8 == 0 && Nmax <= 65535 && plane < (1ull << 31);
  p->native16 = p->fast && label_bytes == 2;
  if (p->fast && p->illum != nullptr) {
    if (ips_illum_reciprocal(p->illum, p->illum, (int64_t)C * (int64_t)plane, nullptr) != IPS_OK) {
      pipeline_free(p);
      return IPS_ERR_CUDA;
    }
    PIPE_CUDA_OK(cudaDeviceSynchronize());
  }
  p->ws_bytes = ips_field_fused_workspace_bytes(fields_per_batch, C, H, W, bin, Nmax);
  p->slots.resize(depth);
  for (Slot& s : p->slots) {
    PIPE_CUDA_OK(cudaMalloc(&s.raw, Fb * C * Z * plane * sizeof(uint16_t)));
    if (!p->native16) PIPE_CUDA_OK(cudaMalloc(&s.labels, Fb * plane * sizeof(int32_t)));
    if (label_bytes == 2) PIPE_CUDA_OK(cudaMalloc(&s.labels16, Fb * plane * sizeof(uint16_t)));
    PIPE_CUDA_OK(cudaMalloc(&s.maxproj, Fb * C * plane * sizeof(uint16_t)));
    PIPE_CUDA_OK(cudaMalloc(&s.binned, Fb * C * (plane / (bin * bin)) * 4));
    PIPE_CUDA_OK(cudaMalloc(&s.n_objects, Fb * sizeof(int32_t)));
    PIPE_CUDA_OK(cudaMalloc(&s.ints, Fb * Nmax * 6 * sizeof(int32_t)));
    PIPE_CUDA_OK(cudaMalloc(&s.flts, Fb * Nmax * (2 + 5 * (size_t)C) * sizeof(float)));
    PIPE_CUDA_OK(cudaMalloc(&s.ws, p->ws_bytes));
    PIPE_CUDA_OK(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
    PIPE_CUDA_OK(cudaEventCreateWithFlags(&s.compute_done, cudaEventDisableTiming));
    PIPE_CUDA_OK(cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming));
  }
  *out = p;
  return IPS_OK;
}

extern "C" int64_t ips_pipeline_submit(ips_pipeline_t* p, const uint16_t* raw_host,
                                       const void* labels_host, uint16_t* maxproj_host,
                                       void* binned_host, int32_t* n_objects_host,
                                       int32_t* ints_host, float* flts_host) {
  if (p == nullptr || raw_host == nullptr || labels_host == nullptr)
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_pipeline_submit: NULL pipeline or input");
  std::lock_guard<std::mutex> lock(p->mu);
  const int64_t ticket = p->next_ticket;
  Slot& s = p->slots[ticket % p->depth];
  const size_t plane = (size_t)p->H * p->W, Fb = (size_t)p->Fb, C = (size_t)p->C;
  // The slot's previous occupant must have left the device before it is overwritten.
  if (s.ticket >= 0) IPS_CUDA_OK(cudaStreamWaitEvent(p->s_in, s.d2h_done, 0));
  IPS_CUDA_OK(cudaMemcpyAsync(s.raw, raw_host, Fb * C * p->Z * plane * sizeof(uint16_t),
                              cudaMemcpyHostToDevice, p->s_in));
  if (p->label_bytes == 4)
    IPS_CUDA_OK(cudaMemcpyAsync(s.labels, labels_host, Fb * plane * sizeof(int32_t), cudaMemcpyHostToDevice, p->s_in));
  else
    IPS_CUDA_OK(cudaMemcpyAsync(s.labels16, labels_host, Fb * plane * sizeof(uint16_t), cudaMemcpyHostToDevice, p->s_in));
  IPS_CUDA_OK(cudaEventRecord(s.h2d_done, p->s_in));
  IPS_CUDA_OK(cudaStreamWaitEvent(p->s_compute, s.h2d_done, 0));
  if (p->label_bytes == 2 && !p->native16) {
    const size_t n = Fb * plane, words = n / 8;
    widen_labels_kernel<<<(unsigned)((words + 256) / 256), 256, 0, p->s_compute>>>(s.labels16, s.labels, words, n);
    IPS_LAUNCH_OK("widen_labels_kernel");
  }
  int rc;
  if (p->fast)
    rc = ips_field_fused_ex(s.raw, p->illum, p->illum != nullptr, p->native16 ? (const void*)s.labels16 : (const void*)s.labels,
                            p->native16 ? 2 : 4, s.maxproj, s.binned, p->bin, p->scale, s.n_objects, s.ints, s.flts,
                            p->Nmax, s.ws, p->ws_bytes, p->Fb, p->C, p->Z, p->H, p->W, p->s_compute);
  else
    rc = ips_field_fused(s.raw, p->illum, s.labels, s.maxproj, s.binned, p->bin, p->scale, s.n_objects, s.ints,
                         s.flts, p->Nmax, s.ws, p->ws_bytes, p->Fb, p->C, p->Z, p->H, p->W, p->s_compute);
  if (rc != IPS_OK) return rc;
  IPS_CUDA_OK(cudaEventRecord(s.compute_done, p->s_compute));
  IPS_CUDA_OK(cudaStreamWaitEvent(p->s_out, s.compute_done, 0));
  if (maxproj_host)
    IPS_CUDA_OK(cudaMemcpyAsync(maxproj_host, s.maxproj, Fb * C * plane * sizeof(uint16_t),
                                cudaMemcpyDeviceToHost, p->s_out));
  if (binned_host)
    IPS_CUDA_OK(cudaMemcpyAsync(binned_host, s.binned, Fb * C * (plane / (p->bin * p->bin)) * 4,
                                cudaMemcpyDeviceToHost, p->s_out));
  if (n_objects_host)
    IPS_CUDA_OK(cudaMemcpyAsync(n_objects_host, s.n_objects, Fb * sizeof(int32_t), cudaMemcpyDeviceToHost, p->s_out));
  if (ints_host)
    IPS_CUDA_OK(cudaMemcpyAsync(ints_host, s.ints, Fb * p->Nmax * 6 * sizeof(int32_t), cudaMemcpyDeviceToHost, p->s_out));
  if (flts_host)
    IPS_CUDA_OK(cudaMemcpyAsync(flts_host, s.flts, Fb * p->Nmax * (2 + 5 * C) * sizeof(float),
                                cudaMemcpyDeviceToHost, p->s_out));
  IPS_CUDA_OK(cudaEventRecord(s.d2h_done, p->s_out));
  s.ticket = ticket;
  p->next_ticket = ticket + 1;
  return ticket;
}

extern "C" int ips_pipeline_wait(ips_pipeline_t* p, int64_t ticket) {
  if (p == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_pipeline_wait: NULL pipeline");
  cudaEvent_t ev;
  {
    std::lock_guard<std::mutex> lock(p->mu);
    if (ticket < 0 || ticket >= p->next_ticket)
      IPS_FAIL(IPS_ERR_BAD_ARG, "ips_pipeline_wait: unknown ticket %lld", (long long)ticket);
    // If the slot has been reused since, its event was re-recorded behind the newer batch;
    // waiting for that one is a superset of waiting for `ticket` (streams are in order).
    ev = p->slots[ticket % p->depth].d2h_done;
  }
  IPS_CUDA_OK(cudaEventSynchronize(ev));
  return IPS_OK;
}

extern "C" int ips_pipeline_destroy(ips_pipeline_t* p) {
  if (p == nullptr) return IPS_OK;
  cudaError_t e = cudaStreamSynchronize(p->s_in);
  if (e == cudaSuccess) e = cudaStreamSynchronize(p->s_compute);
  if (e == cudaSuccess) e = cudaStreamSynchronize(p->s_out);
  pipeline_free(p);
  if (e != cudaSuccess) IPS_FAIL(IPS_ERR_CUDA, "ips_pipeline_destroy: %s", cudaGetErrorString(e));
  return IPS_OK;
}
