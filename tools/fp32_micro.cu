// Experiment tool: FP32 pipe throughput for FADD / FMUL / FFMA streams with constant vs. all-register operands.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(256) k(float* out, int iters, float c, float d) {
  float r[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) r[i] = threadIdx.x * 0.001f + i;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        if (OP == 0) r[i] = r[i] + c;                                  // FADD reg,const
        if (OP == 1) r[i] = r[i] * c;                                  // FMUL reg,const
        if (OP == 2) r[i] = fmaf(r[i], c, d);                          // FFMA reg,const,const
        if (OP == 3) r[i] = r[i] + r[(i + 5) % 12];                    // FADD reg,reg
        if (OP == 4) r[i] = r[i] * r[(i + 5) % 12];                    // FMUL reg,reg
        if (OP == 5) r[i] = fmaf(r[i], r[(i + 5) % 12], r[(i + 7) % 12]);  // FFMA 3 distinct regs
        if (OP == 6) r[i] = fmaf(r[(i + 5) % 12], r[(i + 5) % 12], r[i]);  // FFMA a*a+c
        if (OP == 7) { r[i] = r[i] + r[(i + 5) % 12]; r[(i + 1) % 12] = fmaf(r[(i + 1) % 12], c, d); }  // FADD + FFMA pairs
        if (OP == 8) { r[i] = r[i] * r[(i + 5) % 12]; r[(i + 1) % 12] = fmaf(r[(i + 1) % 12], c, d); }  // FMUL + FFMA pairs
      }
  }
  float s = 0;
  for (int i = 0; i < 12; ++i) s += r[i];
  if (s == 1234.5f) out[0] = s;
}
template <int OP>
void run(const char* nm, float* out, int sms, double clk, int per) {
  int blocks = sms * 8, iters = 2000;
  k<OP><<<blocks, 256>>>(out, iters, 1.0000001f, 1e-9f);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); k<OP><<<blocks, 256>>>(out, iters, 1.0000001f, 1e-9f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double winstr = double(blocks) * 8 * iters * 48 * per;  // warp instructions
  double cyc = ms * 1e-3 * clk * sms * 4;                 // SMSP cycles available
  printf("%-28s %7.3f ms  %.3f cycles per warp instruction per SMSP\n", nm, ms, cyc / winstr);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double clk = p.clockRate * 1e3; int sms = p.multiProcessorCount;
  float* out; cudaMalloc(&out, 64);
  run<0>("FADD r,c", out, sms, clk, 1);
  run<1>("FMUL r,c", out, sms, clk, 1);
  run<2>("FFMA r,c,c", out, sms, clk, 1);
  run<3>("FADD r,r", out, sms, clk, 1);
  run<4>("FMUL r,r", out, sms, clk, 1);
  run<5>("FFMA r,r,r (3 distinct)", out, sms, clk, 1);
  run<6>("FFMA a,a,r", out, sms, clk, 1);
  run<7>("FADD r,r + FFMA r,c,c", out, sms, clk, 2);
  run<8>("FMUL r,r + FFMA r,c,c", out, sms, clk, 2);
  return 0;
}
