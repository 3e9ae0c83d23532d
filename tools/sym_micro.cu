// Experiment tool: inner loop of a symmetric (Newton's third law) all-pairs tile: each lane holds RI i-bodies, the warp
// sweeps j bodies from shared memory; per pair BOTH a_i += m_j*d*inv and r_j -= m_i*d*inv are accumulated; the reaction
// partial r_j is reduced over the warp with a transposed butterfly every JB bodies and stored to smem per warp.
#include <cfloat>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float msqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mrcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr int TILE = 256;

template <int RI, int MINB>
__global__ void __launch_bounds__(256, MINB) sym_kernel(const float4* __restrict__ src, float4* out, int reps) {
  __shared__ float4 tile[TILE];
  __shared__ float racc[8][TILE][3];
  for (int q = threadIdx.x; q < TILE; q += 256) tile[q] = src[q];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float xi[RI], yi[RI], zi[RI], mi[RI], ax[RI], ay[RI], az[RI];
#pragma unroll
  for (int t = 0; t < RI; ++t) {
    float4 b = src[(blockIdx.x * 256 + threadIdx.x + t * 97) % TILE];
    xi[t] = b.x + 0.37f; yi[t] = b.y - 0.11f; zi[t] = b.z + 0.05f; mi[t] = b.w;
    ax[t] = ay[t] = az[t] = 0.f;
  }
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
#pragma unroll 1
    for (int j0 = 0; j0 < TILE; j0 += 4) {
      float rx[4], ry[4], rz[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 b = tile[j0 + jj];
        rx[jj] = ry[jj] = rz[jj] = 0.f;
#pragma unroll
        for (int t = 0; t < RI; ++t) {
          float dx = b.x - xi[t], dy = b.y - yi[t], dz = b.z - zi[t];
          float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
          float inv = mrcp(fmaf(d2, msqrt(d2), FLT_EPSILON));
          float si = b.w * inv, sj = mi[t] * inv;
          ax[t] = fmaf(dx, si, ax[t]); ay[t] = fmaf(dy, si, ay[t]); az[t] = fmaf(dz, si, az[t]);
          rx[jj] = fmaf(-dx, sj, rx[jj]); ry[jj] = fmaf(-dy, sj, ry[jj]); rz[jj] = fmaf(-dz, sj, rz[jj]);
        }
      }
      // transposed butterfly: 12 values -> lanes; after 2 halving stages each lane holds 3 values, then 3 full stages
      float v[12] = {rx[0], ry[0], rz[0], rx[1], ry[1], rz[1], rx[2], ry[2], rz[2], rx[3], ry[3], rz[3]};
      // stage xor 16: lanes <16 keep first 6, others keep last 6
      float w[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        float mine = (lane & 16) ? v[6 + q] : v[q];
        float send = (lane & 16) ? v[q] : v[6 + q];
        w[q] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
      }
      float u[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        float mine = (lane & 8) ? w[3 + q] : w[q];
        float send = (lane & 8) ? w[q] : w[3 + q];
        u[q] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        u[q] += __shfl_xor_sync(0xffffffffu, u[q], 4);
        u[q] += __shfl_xor_sync(0xffffffffu, u[q], 2);
        u[q] += __shfl_xor_sync(0xffffffffu, u[q], 1);
      }
      if ((lane & 7) == 0) {
        int jsel = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
        racc[warp][j0 + jsel][0] = u[0]; racc[warp][j0 + jsel][1] = u[1]; racc[warp][j0 + jsel][2] = u[2];
      }
    }
  }
  float sx = 0, sy = 0, sz = 0;
#pragma unroll
  for (int t = 0; t < RI; ++t) { sx += ax[t]; sy += ay[t]; sz += az[t]; }
  __syncthreads();
  sx += racc[warp][threadIdx.x % TILE][0];
  out[blockIdx.x * 256 + threadIdx.x] = make_float4(sx, sy, sz, 0);
}

template <int RI, int MINB>
void run(const float4* src, float4* out, int sms, double clk) {
  const int blocks = sms * MINB * 4, reps = 20;
  sym_kernel<RI, MINB><<<blocks, 256>>>(src, out, reps);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); sym_kernel<RI, MINB><<<blocks, 256>>>(src, out, reps); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double upairs = double(blocks) * 256 * RI * TILE * reps;  // unordered pairs evaluated
  double per = ms * 1e-3 * clk / (upairs / 32 / (sms * 4));
  printf("symmetric RI=%d MINB=%d : %7.3f ms  %6.2f cycles per 32 unordered pairs per SMSP = %.2f per 32 ordered  (%.0f G ordered pairs/s)\n", RI,
         MINB, ms, per, per / 2, 2 * upairs / ms / 1e6);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double clk = p.clockRate * 1e3;
  float4 *src, *out;
  cudaMalloc(&src, TILE * sizeof(float4)); cudaMalloc(&out, sizeof(float4) * 148 * 32 * 256);
  float4 h[TILE];
  for (int i = 0; i < TILE; ++i) h[i] = make_float4(i * 0.731f, (i * 37 % 101) * 0.5f, (i * 11 % 53) * 0.25f, 1e-3f);
  cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
  int sms = p.multiProcessorCount;
  run<4, 2>(src, out, sms, clk);
  run<8, 2>(src, out, sms, clk);
  run<8, 1>(src, out, sms, clk);
  run<4, 3>(src, out, sms, clk);
  return 0;
}
