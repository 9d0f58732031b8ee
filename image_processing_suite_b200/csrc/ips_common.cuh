// Shared device helpers and host-side error plumbing for the ips C-ABI library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/ips.h"

namespace ips {

// ---- host: thread-local error string, launch counter -------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define IPS_FAIL(code, ...)            \
  do {                                 \
    ::ips::set_error(__VA_ARGS__);     \
    return (code);                     \
  } while (0)

#define IPS_CUDA_OK(expr)                                                             \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess)                                                           \
      IPS_FAIL(IPS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
               __FILE__, __LINE__);                                                   \
  } while (0)

#define IPS_LAUNCH_OK(name)                                                             \
  do {                                                                                  \
    ::ips::count_launch();                                                              \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess)                                                             \
      IPS_FAIL(IPS_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
int sm_count();  // cached

// ---- device: 128-bit global accesses with L2 eviction policies -----------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// streaming read: read-only path, do not allocate in L1, L2 policy given
__device__ __forceinline__ uint4 ldg128_stream(const void* ptr, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(ptr), "l"(pol));
  return r;
}
// re-used read (illumination function): keep in L2
__device__ __forceinline__ uint4 ldg128_keep(const void* ptr, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(ptr), "l"(pol));
  return r;
}
__device__ __forceinline__ void stg128_stream(void* ptr, uint4 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
               :: "l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void stg64_stream(void* ptr, uint2 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;"
               :: "l"(ptr), "r"(v.x), "r"(v.y), "l"(pol)
               : "memory");
}
// x / d as MUFU.RCP + FMUL (about 2 ulp).  __fdividef adds a range-scaling prologue of three
// more instructions per quotient that the streaming kernels cannot afford; illumination
// functions are O(1), far from the denormal / 2^126 ranges that prologue protects.
__device__ __forceinline__ float fast_div(float x, float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return x * r;
}
// max of packed uint16 pairs
__device__ __forceinline__ uint32_t vmax_u16x2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
__device__ __forceinline__ uint4 vmax_u16x8(uint4 a, uint4 b) {
  return make_uint4(__vmaxu2(a.x, b.x), __vmaxu2(a.y, b.y), __vmaxu2(a.z, b.z), __vmaxu2(a.w, b.w));
}
__device__ __forceinline__ void unpack_u16x8(uint4 v, float out[8]) {
  out[0] = (float)(v.x & 0xFFFFu); out[1] = (float)(v.x >> 16);
  out[2] = (float)(v.y & 0xFFFFu); out[3] = (float)(v.y >> 16);
  out[4] = (float)(v.z & 0xFFFFu); out[5] = (float)(v.z >> 16);
  out[6] = (float)(v.w & 0xFFFFu); out[7] = (float)(v.w >> 16);
}
__device__ __forceinline__ void unpack_u16x8(uint4 v, uint32_t out[8]) {
  out[0] = v.x & 0xFFFFu; out[1] = v.x >> 16;
  out[2] = v.y & 0xFFFFu; out[3] = v.y >> 16;
  out[4] = v.z & 0xFFFFu; out[5] = v.z >> 16;
  out[6] = v.w & 0xFFFFu; out[7] = v.w >> 16;
}
#endif  // __CUDACC__

}  // namespace ips
