// Microbenchmark (experiment tool, not product): throughput of MUFU ops and a few FP32 instruction mixes on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 1) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 2) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 3) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 4) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 5) asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 6) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int OP>
__global__ void __launch_bounds__(256) mufu_kernel(float* out, int iters, float seed) {
  float r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = seed + threadIdx.x * 1e-3f + k;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = op<OP>(r[k]);
  }
  float s = 0;
  for (int k = 0; k < 8; ++k) s += r[k];
  if (s == 12345.678f) out[0] = s;
}

// rcp + sqrt interleaved (what the all-pairs kernel issues)
__global__ void __launch_bounds__(256) mix_kernel(float* out, int iters, float seed) {
  float r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = seed + threadIdx.x * 1e-3f + k;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = op<0>(op<1>(r[k]));
  }
  float s = 0;
  for (int k = 0; k < 8; ++k) s += r[k];
  if (s == 12345.678f) out[0] = s;
}

template <typename F>
double time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* out; cudaMalloc(&out, 64);
  const int blocks = p.multiProcessorCount * 8, iters = 4000;
  const double ops = double(blocks) * 256 * iters * 32;
  const char* names[] = {"rcp", "sqrt", "rsqrt", "ex2", "lg2", "sin", "tanh"};
  double clk = p.clockRate * 1e3;
  auto report = [&](const char* nm, double ms) {
    double per_clk_sm = ops / (ms * 1e-3) / clk / p.multiProcessorCount;
    printf("%-8s %8.3f ms  %6.2f lanes/clk/SM  (%.2f cycles per warp instr per SMSP)\n", nm, ms, per_clk_sm, 32.0 * 4 / per_clk_sm);
  };
  report(names[0], time_ms([&] { mufu_kernel<0><<<blocks, 256>>>(out, iters, 1.5f); }));
  report(names[1], time_ms([&] { mufu_kernel<1><<<blocks, 256>>>(out, iters, 1.5f); }));
  report(names[2], time_ms([&] { mufu_kernel<2><<<blocks, 256>>>(out, iters, 1.5f); }));
  report(names[3], time_ms([&] { mufu_kernel<3><<<blocks, 256>>>(out, iters, 0.5f); }));
  report(names[4], time_ms([&] { mufu_kernel<4><<<blocks, 256>>>(out, iters, 1.5f); }));
  report(names[5], time_ms([&] { mufu_kernel<5><<<blocks, 256>>>(out, iters, 1.5f); }));
  report(names[6], time_ms([&] { mufu_kernel<6><<<blocks, 256>>>(out, iters, 1.5f); }));
  report("sqrt+rcp", time_ms([&] { mix_kernel<<<blocks, 256>>>(out, iters, 1.5f); }));
  printf("clock %.0f MHz, %d SMs\n", clk / 1e6, p.multiProcessorCount);
  return 0;
}
