// Experiment tool: accuracy of rsqrt.approx.ftz.f64 / rcp.approx.ftz.f64 seeds and of the refinement formulas used in libnbx.
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = exp2(-60.0 + 120.0 * (double(i) + 0.37) / n);  // 2^-60 .. 2^60
  double y, r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double ey = fabs(y * sqrt(x) - 1.0), er = fabs(r * x - 1.0);
  // cubic refinement of rsqrt
  double t = x * y, e = fma(-t, y, 1.0), q = e * fma(0.375, e, 0.5), y1 = fma(y, q, y);
  double sq = x * y1;
  double es = fabs(sq / sqrt(x) - 1.0);
  // cubic refinement of rcp
  double e1 = fma(-x, r, 1.0), r1 = fma(r, fma(e1, e1, e1), r);
  double er1 = fabs(r1 * x - 1.0);
  out[4 * i] = ey; out[4 * i + 1] = er; out[4 * i + 2] = es; out[4 * i + 3] = er1;
}
int main() {
  int n = 1 << 22;
  double* d; cudaMalloc(&d, sizeof(double) * 4 * n);
  k<<<(n + 255) / 256, 256>>>(d, n);
  double* h = new double[4 * n];
  cudaMemcpy(h, d, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost);
  double m[4] = {0, 0, 0, 0};
  for (int i = 0; i < n; ++i) for (int c = 0; c < 4; ++c) if (h[4 * i + c] > m[c]) m[c] = h[4 * i + c];
  printf("max rel err: rsqrt.approx.f64 %.3e (2^%.1f)  rcp.approx.f64 %.3e (2^%.1f)  sqrt after cubic %.3e  rcp after cubic %.3e\n", m[0], log2(m[0]), m[1], log2(m[1]), m[2], m[3]);
  return 0;
}
