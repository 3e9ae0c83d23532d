// Experiment tool: how many issue/dispatch cycles does a MUFU cost next to K FMA-pipe instructions?
#include <cstdio>
#include <cuda_runtime.h>
template <int K, int M>
__global__ void __launch_bounds__(256) k(float* out, int iters, float c, float d) {
  float r[16], m[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = 1.5f + threadIdx.x * 0.01f + i;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int q = 0; q < M; ++q) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(m[(u + q) % 4]));
#pragma unroll
      for (int i = 0; i < K; ++i) r[(u * K + i) % 16] = fmaf(r[(u * K + i) % 16], c, d);
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += r[i];
  for (int i = 0; i < 4; ++i) s += m[i];
  if (s == 1234.5f) out[0] = s;
}
template <int K, int M>
void run(float* out, int sms, double clk) {
  int blocks = sms * 8, iters = 4000;
  k<K, M><<<blocks, 256>>>(out, iters, 1.0000001f, 1e-9f);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); k<K, M><<<blocks, 256>>>(out, iters, 1.0000001f, 1e-9f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double groups = double(blocks) * 8 * iters * 4;  // warp-level (M MUFU + K FFMA) groups
  double cyc = ms * 1e-3 * clk * sms * 4;
  printf("M=%d MUFU + K=%2d FFMA: %.2f cycles per group per SMSP\n", M, K, cyc / groups);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double clk = p.clockRate * 1e3; int sms = p.multiProcessorCount;
  float* out; cudaMalloc(&out, 64);
  run<0, 1>(out, sms, clk); run<2, 1>(out, sms, clk); run<4, 1>(out, sms, clk); run<6, 1>(out, sms, clk); run<8, 1>(out, sms, clk);
  run<10, 1>(out, sms, clk); run<12, 1>(out, sms, clk); run<16, 1>(out, sms, clk);
  run<8, 2>(out, sms, clk); run<11, 2>(out, sms, clk); run<14, 2>(out, sms, clk); run<16, 2>(out, sms, clk);
  run<8, 0>(out, sms, clk); run<16, 0>(out, sms, clk);
  return 0;
}
