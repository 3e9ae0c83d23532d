// Experiment tool: does sm_100a have packed FP32x2 FMA (fma.rn.f32x2 -> SASS FFMA2) and what is its throughput?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
template <int OP>
__global__ void __launch_bounds__(256) k(unsigned long long* out, int iters, unsigned long long c, unsigned long long d) {
  unsigned long long r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = (unsigned long long)__float_as_uint(threadIdx.x * 0.001f + i) * 0x100000001ull;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == 0) r[i] = ffma2(r[i], c, d);
        if (OP == 1) r[i] = ffma2(r[i], r[(i + 3) % 8], r[(i + 5) % 8]);
        if (OP == 2) r[i] = fadd2(r[i], r[(i + 3) % 8]);
        if (OP == 3) r[i] = fmul2(r[i], r[(i + 3) % 8]);
      }
  }
  unsigned long long s = 0;
  for (int i = 0; i < 8; ++i) s ^= r[i];
  if (s == 12345ull) out[0] = s;
}
template <int OP>
void run(const char* nm, unsigned long long* out, int sms, double clk) {
  int blocks = sms * 8, iters = 4000;
  float one = 1.0000001f, eps = 1e-9f;
  unsigned long long c = (unsigned long long)(*(unsigned*)&one) * 0x100000001ull, d = (unsigned long long)(*(unsigned*)&eps) * 0x100000001ull;
  k<OP><<<blocks, 256>>>(out, iters, c, d);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); k<OP><<<blocks, 256>>>(out, iters, c, d); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double winstr = double(blocks) * 8 * iters * 32;
  double cyc = ms * 1e-3 * clk * sms * 4;
  printf("%-28s %7.3f ms  %.3f cycles per warp instruction per SMSP  (%.1f TFLOP/s-equivalent)\n", nm, ms, cyc / winstr,
         winstr * 32 * 2 * (OP <= 1 ? 2 : 1) / (ms * 1e-3) / 1e12);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double clk = p.clockRate * 1e3; int sms = p.multiProcessorCount;
  unsigned long long* out; cudaMalloc(&out, 64);
  run<0>("FFMA2 r,c,c", out, sms, clk);
  run<1>("FFMA2 r,r,r", out, sms, clk);
  run<2>("FADD2 r,r", out, sms, clk);
  run<3>("FMUL2 r,r", out, sms, clk);
  return 0;
}
