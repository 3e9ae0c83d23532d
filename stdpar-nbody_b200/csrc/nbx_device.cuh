// nbx_device.cuh — device helpers shared by the all-pairs translation units: exact (non-contracted) arithmetic, TMA bulk
// copy + mbarrier PTX wrappers, L2-only vec4 accesses and the leapfrog update.
#pragma once
#include "nbx_internal.cuh"
#include "nbx_math.cuh"

namespace nbx {

// (mul_rn / add_rn: nbx_math.cuh. The leapfrog restates system.h:56-58 operation by operation so that, given the same
// `a`, it is bit-identical to the pinned reference.)

// ---- TMA bulk copy + mbarrier helpers (PTX) -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D bulk async copy global -> shared, completion signalled on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2-only (cache-global) vec4 accesses for data produced by other CTAs of the same launch
__device__ __forceinline__ float4 ldcg_v4(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ double4 ldcg_v4(const double4* p) {
  double2 lo = __ldcg(reinterpret_cast<const double2*>(p));
  double2 hi = __ldcg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void stcg_v4(float4* p, float4 v) { __stcg(p, v); }
__device__ __forceinline__ void stcg_v4(double4* p, double4 v) {
  __stcg(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
  __stcg(reinterpret_cast<double2*>(p) + 1, make_double2(v.z, v.w));
}

// ---- leapfrog (system.h:52-60) --------------------------------------------------------------------------------
template <typename T>
struct LeapArgs {
  const vec4_t<T>* xm_in;
  vec4_t<T>* xm_out;
  vec4_t<T>* v;
  vec4_t<T>* a;
  vec4_t<T>* ao;
  T dt;
};

// x += dt*v + 0.5*dt*dt*ao ; v += 0.5*dt*(a+ao) ; ao = a     (all vec ops componentwise, same association)
template <typename T, int D>
__device__ __forceinline__ void leapfrog_body(const LeapArgs<T>& p, uint32_t i, vec4_t<T> anew) {
  vec4_t<T> xm = p.xm_in[i];
  vec4_t<T> v  = p.v[i];
  vec4_t<T> ao = p.ao[i];
  const T hdt2 = mul_rn(mul_rn(T(0.5), p.dt), p.dt);
  const T hdt  = mul_rn(T(0.5), p.dt);
  xm.x = add_rn(xm.x, add_rn(mul_rn(v.x, p.dt), mul_rn(ao.x, hdt2)));
  xm.y = add_rn(xm.y, add_rn(mul_rn(v.y, p.dt), mul_rn(ao.y, hdt2)));
  if (D == 3) xm.z = add_rn(xm.z, add_rn(mul_rn(v.z, p.dt), mul_rn(ao.z, hdt2)));
  v.x = add_rn(v.x, mul_rn(add_rn(anew.x, ao.x), hdt));
  v.y = add_rn(v.y, mul_rn(add_rn(anew.y, ao.y), hdt));
  if (D == 3) v.z = add_rn(v.z, mul_rn(add_rn(anew.z, ao.z), hdt));
  p.xm_out[i] = xm;
  p.v[i]      = v;
  p.ao[i]     = anew;
}


}  // namespace nbx
