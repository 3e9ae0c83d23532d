// nbx_math.cuh — the softened inverse-distance kernels shared by the force kernels (sm_100a).
//
// Seeds come from the XU pipe (MUFU.SQRT / MUFU.RCP for float, MUFU.RSQ64H / MUFU.RCP64H for double, measured on B200:
// max relative error 2^-20.1 / 2^-20.0 for the double seeds); double results are refined on the FP64 pipe with ONE cubic
// step each (error^3 => 2^-60, below the rounding of the surrounding operations; measured 1-2 ulp end to end).
#pragma once
#include <cfloat>

namespace nbx {

// ---- exact arithmetic ---------------------------------------------------------------------------------------------------
// Round-to-nearest operations that the compiler may not contract into FMAs: everything that feeds an integer artefact, a
// tree node, an accept/open decision or the leapfrog restates the reference's expressions operation by operation with
// these, so that the results are bit-identical to the pinned (-ffp-contract=off) reference.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// ---- 1 / (d2^1.5 + eps)   (vec.h:249-252 dist3; all-pairs and bvh) ---------------------------------------------------
// d2 == 0 gives 1/eps (finite), so a self / coincident pair contributes m * 0 * (1/eps) = 0 exactly as the reference's
// m*(pj-pi)/dist3 does, without a branch.
__device__ __forceinline__ float inv_dist3(float d2) {
  float sq, inv;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(d2));  // MUFU.SQRT
  float den = fmaf(d2, sq, FLT_EPSILON);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(den));  // MUFU.RCP
  return inv;
}

// sqrt(d2) in double: rsqrt seed + cubic step  y1 = y0 (1 + e/2 + 3e^2/8), e = 1 - d2 y0^2 ; sqrt = d2 * y1  (0 -> 0)
__device__ __forceinline__ double sqrt_fast(double d2) {
  const double d2c = fmax(d2, 1e-300);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d2c));  // MUFU.RSQ64H
  const double t = d2c * y;
  const double e = fma(-t, y, 1.0);
  const double q = e * fma(0.375, e, 0.5);
  y              = fma(y, q, y);
  return d2 * y;
}
// 1/x in double (x > 0): rcp seed + cubic step  r1 = r0 (1 + e + e^2), e = 1 - x r0
__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // MUFU.RCP64H
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double inv_dist3(double d2) { return rcp_fast(fma(d2, sqrt_fast(d2), DBL_EPSILON)); }

// All-pairs inner loops: in double d2 is formed as (dx*dx + 1e-300) + dy*dy (+ dz*dz) — the addend rides on the first FMA for free
// and keeps d2 > 0, so the double path needs no fmax guard in front of the rsqrt seed. 1e-300 is far below the rounding
// of any non-zero d2; for a self / coincident pair it gives den = eps exactly as d2 = 0 does, and the pair still
// contributes m * 0 * (1/eps) = 0.
__device__ __forceinline__ float sq_plus_tiny(float dx) { return dx * dx; }  // float: MUFU.SQRT(0) = 0 needs no guard
__device__ __forceinline__ double sq_plus_tiny(double dx) { return fma(dx, dx, 1e-300); }
__device__ __forceinline__ float inv_dist3_pos(float d2) { return inv_dist3(d2); }
__device__ __forceinline__ double inv_dist3_pos(double d2) {  // requires d2 > 0
  // (One seed instead of two — y = d2^-1/2, y^3 (1 - eps y^3), the two-seed form only for d2 < 1e-6 — saves 2 of the 25 DP
  // operations per unordered pair on paper; measured at n = 262 144: 93.3 instead of 56.7 ms per step — the data-dependent
  // branch breaks the register-blocked loop, as the float one-MUFU trial did. Not kept.)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d2));
  const double t = d2 * y;
  const double e = fma(-t, y, 1.0);
  const double q = e * fma(0.375, e, 0.5);
  y              = fma(y, q, y);
  return rcp_fast(fma(d2, d2 * y, DBL_EPSILON));
}

// ---- octree form: dx = sqrt(d2) + eps ; 1/dx^3   (vec.h:243-246 dist, octree.h:240-242) ---------------------------------
__device__ __forceinline__ float dist_eps(float d2) {
  float sq;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(d2));
  return sq + FLT_EPSILON;
}
__device__ __forceinline__ double dist_eps(double d2) { return sqrt_fast(d2) + DBL_EPSILON; }
// same with d2 > 0 guaranteed by sq_plus_tiny (sqrt(1e-300) + eps == eps exactly, like the reference's 0 + eps)
__device__ __forceinline__ float dist_eps_pos(float d2) { return dist_eps(d2); }
__device__ __forceinline__ double dist_eps_pos(double d2) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d2));
  const double t = d2 * y;
  const double e = fma(-t, y, 1.0);
  const double q = e * fma(0.375, e, 0.5);
  y              = fma(y, q, y);
  return fma(d2, y, DBL_EPSILON);
}
// 1 / (sqrt(d2) + eps) for the energies (vec.h:243-246 dist): MUFU seeds; double refined like the force kernels.
// (A Newton step on each float seed was measured: 1.5x slower, no change in the summed energy — the float error is in the
// accumulation, not in the seeds.)
__device__ __forceinline__ float inv_dist_eps(float d2) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dist_eps(d2)));
  return r;
}
__device__ __forceinline__ double inv_dist_eps(double d2) { return rcp_fast(dist_eps_pos(d2)); }  // d2 > 0 (sq_plus_tiny)
__device__ __forceinline__ float inv_cube(float dx) {
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(dx * dx * dx));
  return inv;
}
__device__ __forceinline__ double inv_cube(double dx) { return rcp_fast(dx * dx * dx); }

}  // namespace nbx
