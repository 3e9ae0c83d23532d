// Experiment tool: the symmetric all-pairs inner loop of sym_micro.cu with PACKED FP32x2 arithmetic (sm_100a FFMA2 / FADD2 /
// FMUL2): two j bodies per instruction, the j tile in shared memory as SoA (x[], y[], z[], m[]) so that a j pair is one
// LDS.64 per component; the i bodies are held as duplicated pairs. Same FMA-pipe work as the scalar loop, about half the
// issue slots: does the loop move from dispatch-bound towards the XU / FMA-pipe bounds?
#include <cfloat>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float msqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mrcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr int TILE = 256;

template <int RI, int MINB>
__global__ void __launch_bounds__(256, MINB) sym2_kernel(const float4* __restrict__ src, float4* out, int reps) {
  __shared__ __align__(16) float tx[TILE], ty[TILE], tz[TILE], tm[TILE];
  __shared__ float racc[8][TILE][3];
  for (int q = threadIdx.x; q < TILE; q += 256) { float4 b = src[q]; tx[q] = b.x; ty[q] = b.y; tz[q] = b.z; tm[q] = b.w; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2 nxi[RI], nyi[RI], nzi[RI], mi[RI], ax[RI], ay[RI], az[RI];
#pragma unroll
  for (int t = 0; t < RI; ++t) {
    float4 b = src[(blockIdx.x * 256 + threadIdx.x + t * 97) % TILE];
    nxi[t] = make_float2(-(b.x + 0.37f), -(b.x + 0.37f)); nyi[t] = make_float2(-(b.y - 0.11f), -(b.y - 0.11f));
    nzi[t] = make_float2(-(b.z + 0.05f), -(b.z + 0.05f)); mi[t] = make_float2(b.w, b.w);
    ax[t] = ay[t] = az[t] = make_float2(0.f, 0.f);
  }
  const float2 eps2 = make_float2(FLT_EPSILON, FLT_EPSILON);
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
#pragma unroll 1
    for (int j0 = 0; j0 < TILE; j0 += 4) {
      float2 rx[2], ry[2], rz[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const float2 bx = *reinterpret_cast<const float2*>(&tx[j0 + 2 * p]);
        const float2 by = *reinterpret_cast<const float2*>(&ty[j0 + 2 * p]);
        const float2 bz = *reinterpret_cast<const float2*>(&tz[j0 + 2 * p]);
        const float2 bm = *reinterpret_cast<const float2*>(&tm[j0 + 2 * p]);
        rx[p] = ry[p] = rz[p] = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < RI; ++t) {
          const float2 dx = __fadd2_rn(bx, nxi[t]), dy = __fadd2_rn(by, nyi[t]), dz = __fadd2_rn(bz, nzi[t]);
          const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
          const float2 sq = make_float2(msqrt(d2.x), msqrt(d2.y));
          const float2 den = __ffma2_rn(d2, sq, eps2);
          const float2 inv = make_float2(mrcp(den.x), mrcp(den.y));
          const float2 si = __fmul2_rn(bm, inv), sj = __fmul2_rn(mi[t], inv);
          ax[t] = __ffma2_rn(dx, si, ax[t]); ay[t] = __ffma2_rn(dy, si, ay[t]); az[t] = __ffma2_rn(dz, si, az[t]);
          rx[p] = __ffma2_rn(dx, sj, rx[p]); ry[p] = __ffma2_rn(dy, sj, ry[p]); rz[p] = __ffma2_rn(dz, sj, rz[p]);
        }
      }
      float v[12] = {-rx[0].x, -ry[0].x, -rz[0].x, -rx[0].y, -ry[0].y, -rz[0].y, -rx[1].x, -ry[1].x, -rz[1].x, -rx[1].y, -ry[1].y, -rz[1].y};
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
  for (int t = 0; t < RI; ++t) { sx += ax[t].x + ax[t].y; sy += ay[t].x + ay[t].y; sz += az[t].x + az[t].y; }
  __syncthreads();
  sx += racc[warp][threadIdx.x % TILE][0];
  out[blockIdx.x * 256 + threadIdx.x] = make_float4(sx, sy, sz, 0);
}

template <int RI, int MINB>
void run(const float4* src, float4* out, int sms, double clk) {
  const int blocks = sms * MINB * 4, reps = 20;
  sym2_kernel<RI, MINB><<<blocks, 256>>>(src, out, reps);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); sym2_kernel<RI, MINB><<<blocks, 256>>>(src, out, reps); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double upairs = double(blocks) * 256 * RI * TILE * reps;  // unordered pairs evaluated
  double per = ms * 1e-3 * clk / (upairs / 32 / (sms * 4));
  printf("packed symmetric RI=%d MINB=%d : %7.3f ms  %6.2f cycles per 32 unordered pairs per SMSP = %.2f per 32 ordered  (%.0f G ordered pairs/s)\n",
         RI, MINB, ms, per, per / 2, 2 * upairs / ms / 1e6);
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
  run<2, 2>(src, out, sms, clk);
  run<4, 2>(src, out, sms, clk);
  run<4, 3>(src, out, sms, clk);
  run<8, 2>(src, out, sms, clk);
  run<4, 4>(src, out, sms, clk);
  return 0;
}
