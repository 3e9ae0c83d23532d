// Experiment tool (not product): instruction-mix microbenchmarks of the all-pairs inner loop on sm_100a.
// Reports cycles per 32 pairs per SMSP for several formulations so that scheduling effects can be separated from
// pipe limits. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ap_micro tools/ap_micro.cu
#include <cfloat>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float msqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mrcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mrsq(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int TILE = 512;

// MODE 0: product math (sqrt + rcp). 1: FMA only (MUFUs replaced by FMULs). 2: rsq + first-order (1 MUFU).
// 3: sqrt+rcp but only on half of the targets, other half rsq path. 4: MUFU only (no accumulate FMAs beyond 1)
template <int TI, int UNROLL, int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) loop_kernel(const float4* __restrict__ src, float4* out, int reps) {
  __shared__ float4 tile[TILE];
  for (int q = threadIdx.x; q < TILE; q += 256) tile[q] = src[q];
  __syncthreads();
  float xi[TI], yi[TI], zi[TI], ax[TI], ay[TI], az[TI];
#pragma unroll
  for (int t = 0; t < TI; ++t) {
    float4 b = src[(blockIdx.x * 256 + threadIdx.x + t * 97) % TILE];
    xi[t] = b.x + 0.37f; yi[t] = b.y - 0.11f; zi[t] = b.z + 0.05f;
    ax[t] = ay[t] = az[t] = 0.f;
  }
  float qmax = 0.f;
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
#pragma unroll UNROLL
    for (int j = 0; j < TILE; ++j) {
      const float4 b = tile[j];
#pragma unroll
      for (int t = 0; t < TI; ++t) {
        float dx = b.x - xi[t], dy = b.y - yi[t], dz = b.z - zi[t];
        float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        float s;
        if (MODE == 0 || (MODE == 3 && t < TI / 2)) {
          s = b.w * mrcp(fmaf(d2, msqrt(d2), FLT_EPSILON));
        } else if (MODE == 1) {
          s = b.w * (fmaf(d2, d2 * 0.37f, FLT_EPSILON) * 1.01f);
        } else if (MODE == 2 || MODE == 3) {
          float rr = mrsq(d2);
          float q  = (rr * rr) * rr;
          qmax     = fmaxf(qmax, q);
          float mq = b.w * q, u = q * FLT_EPSILON;
          s        = fmaf(-u, mq, mq);
        } else {  // MODE 4: two MUFUs, almost no FMA
          s = mrcp(msqrt(dx));
          ax[t] += s;
          continue;
        }
        ax[t] = fmaf(dx, s, ax[t]);
        ay[t] = fmaf(dy, s, ay[t]);
        az[t] = fmaf(dz, s, az[t]);
      }
    }
  }
  float sx = qmax, sy = 0, sz = 0;
#pragma unroll
  for (int t = 0; t < TI; ++t) { sx += ax[t]; sy += ay[t]; sz += az[t]; }
  out[blockIdx.x * 256 + threadIdx.x] = make_float4(sx, sy, sz, 0);
}

template <int TI, int UNROLL, int MODE, int MINB>
void run(const char* name, const float4* src, float4* out, int sms, double clk) {
  const int blocks = sms * MINB * 4, reps = 40;
  loop_kernel<TI, UNROLL, MODE, MINB><<<blocks, 256>>>(src, out, reps);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  loop_kernel<TI, UNROLL, MODE, MINB><<<blocks, 256>>>(src, out, reps);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double pairs = double(blocks) * 256 * TI * TILE * reps;
  double cyc   = ms * 1e-3 * clk;                       // cycles elapsed
  double per   = cyc / (pairs / 32 / (sms * 4));        // cycles per warp-pair-batch per SMSP
  printf("%-34s TI=%d U=%d MINB=%d : %7.3f ms  %6.2f cycles/32 pairs/SMSP  (%.0f Gpairs/s)\n", name, TI, UNROLL, MINB, ms, per, pairs / ms / 1e6);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double clk = p.clockRate * 1e3;
  float4* src; float4* out;
  cudaMalloc(&src, TILE * sizeof(float4));
  cudaMalloc(&out, sizeof(float4) * 148 * 32 * 256);
  float4 h[TILE];
  for (int i = 0; i < TILE; ++i) h[i] = make_float4(i * 0.731f, (i * 37 % 101) * 0.5f, (i * 11 % 53) * 0.25f, 1e-3f);
  cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
  int sms = p.multiProcessorCount;
  run<4, 4, 0, 3>("sqrt+rcp (product)", src, out, sms, clk);
  run<4, 8, 0, 3>("sqrt+rcp (product)", src, out, sms, clk);
  run<4, 2, 0, 3>("sqrt+rcp (product)", src, out, sms, clk);
  run<4, 1, 0, 3>("sqrt+rcp (product)", src, out, sms, clk);
  run<2, 4, 0, 4>("sqrt+rcp (product)", src, out, sms, clk);
  run<8, 2, 0, 2>("sqrt+rcp (product)", src, out, sms, clk);
  run<4, 4, 0, 2>("sqrt+rcp (product)", src, out, sms, clk);
  run<4, 4, 0, 1>("sqrt+rcp (product)", src, out, sms, clk);
  run<4, 4, 1, 3>("FMA only", src, out, sms, clk);
  run<4, 4, 4, 3>("MUFU only", src, out, sms, clk);
  run<4, 4, 2, 3>("rsq first-order (1 MUFU)", src, out, sms, clk);
  run<4, 4, 3, 3>("half sqrt+rcp, half rsq", src, out, sms, clk);
  run<4, 8, 3, 3>("half sqrt+rcp, half rsq", src, out, sms, clk);
  run<8, 2, 3, 2>("half sqrt+rcp, half rsq", src, out, sms, clk);
  return 0;
}
