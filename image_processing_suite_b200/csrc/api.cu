// Library-level entry points: error string, launch counter, device info, pinned memory.
#include <atomic>
#include <mutex>

#include "ips_common.cuh"

namespace ips {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

}  // namespace ips

extern "C" int ips_abi_version(void) { return IPS_ABI_VERSION; }
extern "C" const char* ips_last_error(void) { return ips::g_err; }
extern "C" uint64_t ips_launch_count(void) { return ips::g_launches.load(); }

extern "C" int ips_device_info(int* sm, int* cc_major, int* cc_minor, size_t* l2_bytes) {
  int dev = 0;
  IPS_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  IPS_CUDA_OK(cudaGetDeviceProperties(&p, dev));
  if (sm) *sm = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (l2_bytes) *l2_bytes = (size_t)p.l2CacheSize;
  return IPS_OK;
}

extern "C" int ips_host_alloc(void** out, size_t bytes) {
  if (out == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_host_alloc: out is NULL");
  IPS_CUDA_OK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return IPS_OK;
}

// flags: 1 = portable (usable from every CUDA context), 2 = write-combined (host writes, device
// reads: the staging buffers of a producer that never reads them back).
extern "C" int ips_host_alloc_flags(void** out, size_t bytes, unsigned flags) {
  if (out == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_host_alloc_flags: out is NULL");
  if (flags & ~3u) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_host_alloc_flags: unknown flags %u", flags);
  unsigned f = cudaHostAllocDefault;
  if (flags & 1u) f |= cudaHostAllocPortable;
  if (flags & 2u) f |= cudaHostAllocWriteCombined;
  IPS_CUDA_OK(cudaHostAlloc(out, bytes, f));
  return IPS_OK;
}

extern "C" int ips_host_free(void* p) {
  if (p != nullptr) IPS_CUDA_OK(cudaFreeHost(p));
  return IPS_OK;
}

// Plain cudaMalloc / cudaFree: allocations of their own (not carved out of a caching allocator's
// segment), as CUDA IPC export needs them (ips_ipc_export, plate.PeerPusher).
extern "C" int ips_device_alloc(void** out, size_t bytes) {
  if (out == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_device_alloc: out is NULL");
  cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    IPS_FAIL(e == cudaErrorMemoryAllocation ? IPS_ERR_NOMEM : IPS_ERR_CUDA, "ips_device_alloc(%zu): %s", bytes, cudaGetErrorString(e));
  }
  return IPS_OK;
}

extern "C" int ips_device_free(void* p) {
  if (p != nullptr) IPS_CUDA_OK(cudaFree(p));
  return IPS_OK;
}
