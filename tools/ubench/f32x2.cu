// Issue-rate microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), FMNMX3, PRMT, I2F, MUFU.RCP.
// nvcc -gencode arch=compute_100a,code=sm_100a -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  unsigned long long p0, p1, p2, p3, A, B;
  asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(x2), "f"(x3));
  asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(x4), "f"(x5));
  asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(x6), "f"(x7));
  asm("mov.b64 %0, {%1,%1};" : "=l"(A) : "f"(a));
  asm("mov.b64 %0, {%1,%1};" : "=l"(B) : "f"(b));
  unsigned u0 = threadIdx.x, u1 = u0 * 3, u2 = u0 * 5, u3 = u0 * 7;
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
    if (MODE == 0) {  // 8 scalar FFMA
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    } else if (MODE == 1) {  // 4 FFMA2 = 8 fma
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(A), "l"(B));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(A), "l"(B));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(A), "l"(B));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(A), "l"(B));
    } else if (MODE == 2) {  // 8 FFMA2 = 16 fma (same instruction count as mode 0)
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(A), "l"(B));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(A), "l"(B));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(A), "l"(B));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(A), "l"(B));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(B), "l"(A));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(B), "l"(A));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(B), "l"(A));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(B), "l"(A));
    } else if (MODE == 3) {  // 8 three-input min
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x0) : "f"(x1), "f"(a));
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x1) : "f"(x2), "f"(a));
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x2) : "f"(x3), "f"(a));
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x3) : "f"(x4), "f"(a));
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x4) : "f"(x5), "f"(a));
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x5) : "f"(x6), "f"(a));
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x6) : "f"(x7), "f"(a));
      asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x7) : "f"(x0), "f"(a));
    } else if (MODE == 4) {  // 8 PRMT
      u0 = __byte_perm(u0, u1, 0x5410); u1 = __byte_perm(u1, u2, 0x7632); u2 = __byte_perm(u2, u3, 0x5410); u3 = __byte_perm(u3, u0, 0x7632);
      u0 = __byte_perm(u0, u1, 0x1054); u1 = __byte_perm(u1, u2, 0x3276); u2 = __byte_perm(u2, u3, 0x1054); u3 = __byte_perm(u3, u0, 0x3276);
    } else if (MODE == 5) {  // 8 I2F
      x0 = (float)(u0 & 0xffff); x1 = (float)(u1 & 0xffff); x2 = (float)(u2 & 0xffff); x3 = (float)(u3 & 0xffff);
      u0 += (unsigned)x0; u1 += (unsigned)x1; u2 += (unsigned)x2; u3 += (unsigned)x3;
    } else if (MODE == 6) {  // 8 MUFU.RCP
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x1));
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x2)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x3));
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x4)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x5));
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x6)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x7));
    }
  }
  float lo, hi, s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (float)(u0 + u1 + u2 + u3);
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p0)); s += lo + hi;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p1)); s += lo + hi;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p2)); s += lo + hi;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p3)); s += lo + hi;
  if (s == 12345.678f) out[0] = s;
}
template <int MODE>
static void run(const char* name, int per_iter) {
  float* d; cudaMalloc(&d, 4);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  dim3 grid(sms * 8), block(256);
  k<MODE><<<grid, block>>>(d, 1.0001f, 0.5f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, block>>>(d, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = (double)grid.x * 8 * ITERS * per_iter;
  printf("%-28s %8.3f ms  %7.2f G warp-instr/s  = %.2f warp-instr/clk/SM at %d MHz nominal\n", name, ms,
         warp_instr / ms / 1e6, warp_instr / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(d);
}
int main() {
  run<0>("8 x FFMA", 8); run<1>("4 x FFMA2 (8 fma)", 4); run<2>("8 x FFMA2 (16 fma)", 8); run<3>("8 x FMNMX3", 8);
  run<4>("8 x PRMT", 8); run<5>("4 x (LOP+I2F+F2I+IADD)", 16); run<6>("8 x MUFU.RCP", 8);
  return 0;
}
