// nbx_allpairs_sym.cu — all-pairs force with Newton's third law (each unordered pair evaluated once), sm_100a.
//
// The reference's all_pairs_force (src/all_pairs.h:14-27) evaluates every ORDERED pair; its own TODO
// (src/all_pairs.h:41-42) notes that the symmetry of the force pairs halves the work. Every pair term is exactly
// antisymmetric in floating point ((xj-xi) == -(xi-xj), same d2), so evaluating it once and applying it to both bodies
// changes nothing but the summation order. The pair kernel is bound by instruction dispatch (11 FMA-pipe + 2 MUFU per
// ordered pair); sharing dx,dy,dz,d2,sqrt,rcp between the two directions costs 15 FMA-pipe + 2 MUFU per
// UNORDERED pair, i.e. ~1.5x fewer issue cycles per ordered pair. In float the pair arithmetic additionally runs on
// packed FP32x2 instructions (two j bodies per FFMA2/FADD2/FMUL2), which halves the issue slots of the FMA-pipe work.
//
// Decomposition (deterministic, no atomics): bodies are cut into K blocks of B bodies (K <= 256). A CTA owns one unit
// (I, J), I <= J: it sweeps the J block through shared memory (TMA bulk copies, 256-body tiles) once per 256*RI-body
// sub-block of I. "Action" sums (on i in I) live in registers over the whole sweep; "reaction" partials (on j in J) are
// reduced across the warp with a transposed shuffle butterfly every 4 bodies, across the 8 warps through shared memory
// every tile, and accumulated by the CTA — the only writer — into P[I][j]. Actions go to P[J][i]. Diagonal units
// (I == J) evaluate ordered pairs without reaction. Afterwards a[b] = c * sum_K P[K][b] in fixed K order, fused with
// the leapfrog update. Every P entry has exactly one writer and every sum a fixed order => bit-reproducible.
// Multi-GPU: units are dealt round-robin to the ranks, the per-rank sums are combined with one ncclAllReduce of the
// accelerations and every rank integrates all bodies (replicated state, no position exchange).
#include <cfloat>
#include <cstdlib>
#include <string>

#include "nbx_device.cuh"
#include "nbx_internal.cuh"
#include "nbx_math.cuh"

namespace nbx {

namespace {

constexpr int SYM_JT     = 256;  // bodies per shared-memory tile
constexpr int SYM_STAGES = 4;
constexpr int SYM_WARPS  = 8;

template <typename T>
struct SymArgs {
  const vec4_t<T>* xm;
  const float* soa;   // packed float kernel only: [4][soa_stride] = x[], y[], z[], m[] of the same bodies
  uint64_t soa_stride;
  vec4_t<T>* P;       // [units of this rank][2][B]: action sums on the unit's I block, reaction sums on its J block
  uint32_t B, K;
  uint32_t unit_begin, unit_stride;
};

__device__ __forceinline__ void unit_to_blocks(uint32_t u, uint32_t& I, uint32_t& J) {
  // u = J*(J+1)/2 + I, 0 <= I <= J
  uint32_t j = uint32_t((sqrt(8.0 * double(u) + 1.0) - 1.0) * 0.5);
  while (uint64_t(j + 1) * (j + 2) / 2 <= u) ++j;
  while (uint64_t(j) * (j + 1) / 2 > u) --j;
  J = j;
  I = u - uint32_t(uint64_t(j) * (j + 1) / 2);
}

// 1 / (d2^1.5 + eps) for a pair of distances: scalar MUFU.SQRT / MUFU.RCP on the halves, the FMA in between packed.
// (A one-MUFU form — r = rsqrt(d2), r^3 (1 - eps r^3), the exact form behind a rarely taken branch for d2 < 6.3e-3 — trades
// 2 MUFU for 3 packed FMA-pipe instructions + a compare per two pairs; measured at n = 1 M: 416 ms instead of 310 ms per
// step, the FMA pipe and the issue slots are the scarcer resource here. Not kept.)
__device__ __forceinline__ float2 inv_dist3_pair(float2 d2) {
  float2 sq, inv;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq.x) : "f"(d2.x));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq.y) : "f"(d2.y));
  const float2 den = __ffma2_rn(d2, sq, make_float2(FLT_EPSILON, FLT_EPSILON));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(den.y));
  return inv;
}

// transposed butterfly over the warp: v = 3 components of 4 consecutive j bodies (12 values) -> 6 -> 3 per lane, then three
// plain stages; lanes 0, 8, 16, 24 end up with the warp totals of bodies j0+0..j0+3 (fixed order => deterministic) and
// store them to racc[warp][j0 + 0..3][3]
template <typename T>
__device__ __forceinline__ void sym_reduce4(const T (&v)[12], int lane, T* racc, int warp, int j0) {
  T w[6], u[3];
  const bool hi16 = lane & 16, hi8 = lane & 8;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const T mine = hi16 ? v[6 + q] : v[q];
    const T send = hi16 ? v[q] : v[6 + q];
    w[q] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const T mine = hi8 ? w[3 + q] : w[q];
    const T send = hi8 ? w[q] : w[3 + q];
    u[q] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    u[q] += __shfl_xor_sync(0xffffffffu, u[q], 4);
    u[q] += __shfl_xor_sync(0xffffffffu, u[q], 2);
    u[q] += __shfl_xor_sync(0xffffffffu, u[q], 1);
  }
  if ((lane & 7) == 0) {
    const int jsel = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
    T* dst = racc + (size_t(warp) * SYM_JT + j0 + jsel) * 3;
    dst[0] = u[0]; dst[1] = u[1]; dst[2] = u[2];
  }
}

// the CTA is the only writer of P[I][j in J]: sum the 8 warps' partials of body `tid` of the tile and accumulate over the
// sub-blocks of I in order
template <typename T>
__device__ __forceinline__ void sym_flush_reactions(const T* racc, vec4_t<T>* dst, int tid, bool accumulate) {
  T rx = T(0), ry = T(0), rz = T(0);
#pragma unroll
  for (int wq = 0; wq < SYM_WARPS; ++wq) {
    const T* src = racc + (size_t(wq) * SYM_JT + tid) * 3;
    rx += src[0]; ry += src[1]; rz += src[2];
  }
  if (accumulate) {
    const vec4_t<T> old = ldcg_v4(dst);
    rx += old.x; ry += old.y; rz += old.z;
  }
  stcg_v4(dst, make_v4<T>(rx, ry, rz, T(0)));
}

// ---- scalar kernel (double; float behind NBX_SYM_PACKED=0) -----------------------------------------------------------------
template <typename T, int D, int RI, int MINB>
__global__ void __launch_bounds__(256, MINB) all_pairs_sym_kernel(SymArgs<T> p) {
  using V4 = vec4_t<T>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  V4* tiles      = reinterpret_cast<V4*>(smem_raw);
  T* racc        = reinterpret_cast<T*>(smem_raw + size_t(SYM_STAGES) * SYM_JT * sizeof(V4));  // [warp][j][3]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + size_t(SYM_STAGES) * SYM_JT * sizeof(V4) + size_t(SYM_WARPS) * SYM_JT * 3 * sizeof(T));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t I, J;
  unit_to_blocks(p.unit_begin + blockIdx.x * p.unit_stride, I, J);
  const bool diag     = I == J;
  const uint32_t I0   = I * p.B, J0 = J * p.B;
  const uint32_t nsub = p.B / (256 * RI), ntile = p.B / SYM_JT;
  const int total     = int(nsub * ntile);
  constexpr uint32_t TILE_BYTES = SYM_JT * sizeof(V4);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < SYM_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int k) {  // k-th tile of the CTA's (isub, jt) sequence: the J block is re-swept for every sub-block of I
    const int stage = k % SYM_STAGES;
    mbar_expect_tx(&bars[stage], TILE_BYTES);
    tma_load_1d(tiles + size_t(stage) * SYM_JT, p.xm + size_t(J0) + size_t(k % int(ntile)) * SYM_JT, TILE_BYTES, &bars[stage]);
  };
  if (tid == 0)
    for (int k = 0; k < SYM_STAGES && k < total; ++k) issue(k);

  V4* Paction = p.P + uint64_t(2 * blockIdx.x) * p.B;      // actions on i in I caused by block J, indexed by i - I0
  V4* Preact  = p.P + uint64_t(2 * blockIdx.x + 1) * p.B;  // reactions on j in J caused by block I, indexed by j - J0

  int k = 0;
  for (uint32_t isub = 0; isub < nsub; ++isub) {
    T xi[RI], yi[RI], zi[RI], mi[RI], ax[RI], ay[RI], az[RI];
#pragma unroll
    for (int t = 0; t < RI; ++t) {
      const V4 b = p.xm[I0 + isub * (256 * RI) + t * 256 + tid];
      xi[t] = b.x; yi[t] = b.y; zi[t] = b.z; mi[t] = b.w;
      ax[t] = ay[t] = az[t] = T(0);
    }
    for (uint32_t jt = 0; jt < ntile; ++jt, ++k) {
      const int stage = k % SYM_STAGES;
      mbar_wait(&bars[stage], (k / SYM_STAGES) & 1);
      const V4* tile = tiles + size_t(stage) * SYM_JT;
      if (diag) {
#pragma unroll 4
        for (int j = 0; j < SYM_JT; ++j) {
          const V4 b = tile[j];
#pragma unroll
          for (int t = 0; t < RI; ++t) {
            T dx = b.x - xi[t], dy = b.y - yi[t];
            T d2 = fma(dy, dy, sq_plus_tiny(dx));
            T dz = T(0);
            if (D == 3) { dz = b.z - zi[t]; d2 = fma(dz, dz, d2); }
            T s   = b.w * inv_dist3_pos(d2);
            ax[t] = fma(dx, s, ax[t]);
            ay[t] = fma(dy, s, ay[t]);
            if (D == 3) az[t] = fma(dz, s, az[t]);
          }
        }
      } else {
#pragma unroll 1
        for (int j0 = 0; j0 < SYM_JT; j0 += 4) {
          T v[12];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const V4 b = tile[j0 + jj];
            T rx = T(0), ry = T(0), rz = T(0);
#pragma unroll
            for (int t = 0; t < RI; ++t) {
              T dx = b.x - xi[t], dy = b.y - yi[t];
              T d2 = fma(dy, dy, sq_plus_tiny(dx));
              T dz = T(0);
              if (D == 3) { dz = b.z - zi[t]; d2 = fma(dz, dz, d2); }
              const T inv = inv_dist3_pos(d2);
              const T si = b.w * inv, sj = mi[t] * inv;
              ax[t] = fma(dx, si, ax[t]);
              ay[t] = fma(dy, si, ay[t]);
              if (D == 3) az[t] = fma(dz, si, az[t]);
              rx = fma(-dx, sj, rx);
              ry = fma(-dy, sj, ry);
              if (D == 3) rz = fma(-dz, sj, rz);
            }
            v[3 * jj] = rx; v[3 * jj + 1] = ry; v[3 * jj + 2] = rz;
          }
          sym_reduce4<T>(v, lane, racc, warp, j0);
        }
        __syncthreads();
        sym_flush_reactions<T>(racc, Preact + jt * SYM_JT + tid, tid, isub != 0);
      }
      __syncthreads();  // tile (and racc) free again
      if (tid == 0 && k + SYM_STAGES < total) issue(k + SYM_STAGES);
    }
#pragma unroll
    for (int t = 0; t < RI; ++t)
      stcg_v4(Paction + isub * (256 * RI) + t * 256 + tid, make_v4<T>(ax[t], ay[t], az[t], T(0)));
  }
}

// ---- FP32x2 kernel --------------------------------------------------------------------------------------------------
// PACKED (float only): the pair arithmetic runs on FP32x2 instructions (FFMA2 / FADD2 / FMUL2), two j bodies per
// instruction — the same FMA-pipe work in about half the issue slots (tools/sym2_micro.cu: 22.7 instead of 25.0 cycles per
// 32 unordered pairs per SMSP). The j tile is staged as SoA (x[], y[], z[], m[]: four bulk copies per tile from the SoA
// copy of the positions) so that a j pair is one LDS.64 per component; the i bodies are held as duplicated pairs.
template <int D, int RI, int MINB>
__global__ void __launch_bounds__(256, MINB) all_pairs_sym_packed_kernel(SymArgs<float> p) {
  using T = float;
  using V4 = vec4_t<T>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  V4* tiles      = reinterpret_cast<V4*>(smem_raw);
  T* racc        = reinterpret_cast<T*>(smem_raw + size_t(SYM_STAGES) * SYM_JT * sizeof(V4));  // [warp][j][3]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + size_t(SYM_STAGES) * SYM_JT * sizeof(V4) + size_t(SYM_WARPS) * SYM_JT * 3 * sizeof(T));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t I, J;
  unit_to_blocks(p.unit_begin + blockIdx.x * p.unit_stride, I, J);
  const bool diag     = I == J;
  const uint32_t I0   = I * p.B, J0 = J * p.B;
  const uint32_t nsub = p.B / (256 * RI), ntile = p.B / SYM_JT;
  const int total     = int(nsub * ntile);
  constexpr uint32_t TILE_BYTES = SYM_JT * sizeof(V4);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < SYM_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int k) {  // k-th tile of the CTA's (isub, jt) sequence: the J block is re-swept for every sub-block of I
    const int stage = k % SYM_STAGES;
    mbar_expect_tx(&bars[stage], TILE_BYTES);
    const size_t j = size_t(J0) + size_t(k % int(ntile)) * SYM_JT;
    float* dst     = reinterpret_cast<float*>(tiles + size_t(stage) * SYM_JT);  // [4][SYM_JT] floats in the same 4 KB
#pragma unroll
    for (int c = 0; c < 4; ++c) tma_load_1d(dst + c * SYM_JT, p.soa + c * p.soa_stride + j, TILE_BYTES / 4, &bars[stage]);
  };
  if (tid == 0)
    for (int k = 0; k < SYM_STAGES && k < total; ++k) issue(k);

  V4* Paction = p.P + uint64_t(2 * blockIdx.x) * p.B;      // actions on i in I caused by block J, indexed by i - I0
  V4* Preact  = p.P + uint64_t(2 * blockIdx.x + 1) * p.B;  // reactions on j in J caused by block I, indexed by j - J0

  int k = 0;
  for (uint32_t isub = 0; isub < nsub; ++isub) {
    float2 nx2[RI], ny2[RI], nz2[RI], m2[RI];  // (-x_i, -x_i) ... (m_i, m_i)
    float2 ax2[RI], ay2[RI], az2[RI];          // even-j / odd-j partial sums
#pragma unroll
    for (int t = 0; t < RI; ++t) {
      const V4 b = p.xm[I0 + isub * (256 * RI) + t * 256 + tid];
      nx2[t] = make_float2(-b.x, -b.x); ny2[t] = make_float2(-b.y, -b.y); nz2[t] = make_float2(-b.z, -b.z);
      m2[t]  = make_float2(b.w, b.w);
      ax2[t] = ay2[t] = az2[t] = make_float2(0.f, 0.f);
    }
    for (uint32_t jt = 0; jt < ntile; ++jt, ++k) {
      const int stage = k % SYM_STAGES;
      mbar_wait(&bars[stage], (k / SYM_STAGES) & 1);
      const V4* tile = tiles + size_t(stage) * SYM_JT;
      {
        const float* tx = reinterpret_cast<const float*>(tile);
        const float *ty = tx + SYM_JT, *tz = tx + 2 * SYM_JT, *tm = tx + 3 * SYM_JT;
        if (diag) {
#pragma unroll 2
          for (int j = 0; j < SYM_JT; j += 2) {
            const float2 bx = *reinterpret_cast<const float2*>(tx + j), by = *reinterpret_cast<const float2*>(ty + j);
            const float2 bz = *reinterpret_cast<const float2*>(tz + j), bm = *reinterpret_cast<const float2*>(tm + j);
#pragma unroll
            for (int t = 0; t < RI; ++t) {
              const float2 dx = __fadd2_rn(bx, nx2[t]), dy = __fadd2_rn(by, ny2[t]);
              float2 d2 = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
              float2 dz = make_float2(0.f, 0.f);
              if (D == 3) { dz = __fadd2_rn(bz, nz2[t]); d2 = __ffma2_rn(dz, dz, d2); }
              const float2 s = __fmul2_rn(bm, inv_dist3_pair(d2));
              ax2[t] = __ffma2_rn(dx, s, ax2[t]);
              ay2[t] = __ffma2_rn(dy, s, ay2[t]);
              if (D == 3) az2[t] = __ffma2_rn(dz, s, az2[t]);
            }
          }
        } else {
#pragma unroll 1
          for (int j0 = 0; j0 < SYM_JT; j0 += 4) {
            T v[12];
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
              const int j = j0 + 2 * pp;
              const float2 bx = *reinterpret_cast<const float2*>(tx + j), by = *reinterpret_cast<const float2*>(ty + j);
              const float2 bz = *reinterpret_cast<const float2*>(tz + j), bm = *reinterpret_cast<const float2*>(tm + j);
              float2 rx = make_float2(0.f, 0.f), ry = rx, rz = rx;  // + sum dx*sj ; the reaction is its negative
#pragma unroll
              for (int t = 0; t < RI; ++t) {
                const float2 dx = __fadd2_rn(bx, nx2[t]), dy = __fadd2_rn(by, ny2[t]);
                float2 d2 = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
                float2 dz = make_float2(0.f, 0.f);
                if (D == 3) { dz = __fadd2_rn(bz, nz2[t]); d2 = __ffma2_rn(dz, dz, d2); }
                const float2 inv = inv_dist3_pair(d2);
                const float2 si = __fmul2_rn(bm, inv), sj = __fmul2_rn(m2[t], inv);
                ax2[t] = __ffma2_rn(dx, si, ax2[t]);
                ay2[t] = __ffma2_rn(dy, si, ay2[t]);
                if (D == 3) az2[t] = __ffma2_rn(dz, si, az2[t]);
                rx = __ffma2_rn(dx, sj, rx);
                ry = __ffma2_rn(dy, sj, ry);
                if (D == 3) rz = __ffma2_rn(dz, sj, rz);
              }
              v[6 * pp + 0] = -rx.x; v[6 * pp + 1] = -ry.x; v[6 * pp + 2] = -rz.x;
              v[6 * pp + 3] = -rx.y; v[6 * pp + 4] = -ry.y; v[6 * pp + 5] = -rz.y;
            }
            sym_reduce4<T>(v, lane, racc, warp, j0);
          }
          __syncthreads();
          sym_flush_reactions<T>(racc, Preact + jt * SYM_JT + tid, tid, isub != 0);
        }
      }
      __syncthreads();  // tile (and racc) free again
      if (tid == 0 && k + SYM_STAGES < total) issue(k + SYM_STAGES);
    }
#pragma unroll
    for (int t = 0; t < RI; ++t)
      stcg_v4(Paction + isub * (256 * RI) + t * 256 + tid,
              make_v4<T>(ax2[t].x + ax2[t].y, ay2[t].x + ay2[t].y, az2[t].x + az2[t].y, T(0)));
  }
}

// what to do with the summed pair terms (sx, sy, sz) of body b:
//   nc == 0 : all_pairs_force            a = c * sum                         (all_pairs.h:26), optionally + leapfrog
//   nc == 2|3: all_pairs_collapsed_force  a[k] = (a[k] - ao[k]) + c * sum[k]  for k < nc  (all_pairs.h:35-39,47-48): the
//              reference "resets" by subtracting the old acceleration and only touches components 0 and 1 (nc = 2)
template <typename T, int D>
__device__ __forceinline__ void sym_store(const LeapArgs<T>& leap, uint32_t b, T sx, T sy, T sz, T c, int fuse, int nc) {
  if (nc == 0) {
    const vec4_t<T> anew = make_v4<T>(mul_rn(sx, c), mul_rn(sy, c), D == 3 ? mul_rn(sz, c) : T(0), T(0));
    leap.a[b] = anew;
    if (fuse) leapfrog_body<T, D>(leap, b, anew);
    return;
  }
  vec4_t<T> q = leap.a[b];
  const vec4_t<T> o = leap.ao[b];
  q.x = add_rn(add_rn(q.x, -o.x), mul_rn(sx, c));
  q.y = add_rn(add_rn(q.y, -o.y), mul_rn(sy, c));
  if (nc == 3 && D == 3) q.z = add_rn(add_rn(q.z, -o.z), mul_rn(sz, c));
  leap.a[b] = q;
}

// a[b] = c * (sum over the units (k, block of b), k ascending, that THIS rank computed), optionally fused with the leapfrog.
// Unit u = (lo, hi) is the rank's local unit (u - unit_begin) / unit_stride; body b gets the unit's action sums when its
// block is the lower index (or the diagonal), the reaction sums when it is the higher one.
template <typename T, int D>
__global__ void __launch_bounds__(256) sym_reduce_kernel(const vec4_t<T>* __restrict__ P, uint32_t B, uint32_t K, uint32_t n,
                                                         uint32_t unit_begin, uint32_t unit_stride, T c, int finish, int fuse, int nc,
                                                         vec4_t<T>* __restrict__ asum, LeapArgs<T> leap) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= n) return;
  const uint32_t Bb = (blockIdx.x * 256) / B;  // B is a multiple of 256: uniform per CTA
  const uint32_t ob = b - Bb * B;
  T sx = 0, sy = 0, sz = 0;
  for (uint32_t k = 0; k < K; ++k) {
    const uint32_t lo = k < Bb ? k : Bb, hi = k < Bb ? Bb : k;
    const uint32_t u  = uint32_t(uint64_t(hi) * (hi + 1) / 2) + lo;
    if (u < unit_begin || (u - unit_begin) % unit_stride) continue;  // another rank's unit
    const uint64_t lu = (u - unit_begin) / unit_stride;
    const vec4_t<T> q = ldcg_v4(P + (2 * lu + (k < Bb ? 1 : 0)) * B + ob);
    sx += q.x; sy += q.y; sz += q.z;
  }
  if (!finish) {
    asum[b] = make_v4<T>(sx, sy, D == 3 ? sz : T(0), T(0));
    return;
  }
  sym_store<T, D>(leap, b, sx, sy, sz, c, fuse, nc);
}

// after the all-reduce of the unscaled sums: scale by c, (optionally) leapfrog
template <typename T, int D>
__global__ void __launch_bounds__(256) sym_finish_kernel(const vec4_t<T>* __restrict__ asum, uint32_t n, T c, int fuse, int nc,
                                                         LeapArgs<T> leap) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= n) return;
  const vec4_t<T> s = asum[b];
  sym_store<T, D>(leap, b, s.x, s.y, s.z, c, fuse, nc);
}

struct SymState {
  uint32_t B = 0, K = 0;
  uint64_t slab = 0;         // K * B: bodies covered by the blocks (>= n)
  void* P    = nullptr;
  void* asum = nullptr;
  float* soa = nullptr;      // packed float kernel: x[], y[], z[], m[] copy of the current positions, refreshed every step
  uint64_t soa_stride = 0;
};

// (x, y, z, m) records -> x[], y[], z[], m[]
__global__ void __launch_bounds__(256) aos_to_soa_kernel(const float4* __restrict__ xm, uint64_t count, uint64_t stride,
                                                         float* __restrict__ soa) {
  const uint64_t i = blockIdx.x * 256ull + threadIdx.x;
  if (i >= count) return;
  const float4 b = xm[i];
  soa[i] = b.x; soa[stride + i] = b.y; soa[2 * stride + i] = b.z; soa[3 * stride + i] = b.w;
}

}  // namespace

// B: 1024 below n = 2^17 (K = ceil(n / B) <= 128: enough (I, J) units for 148 SMs from n = 16384 on), above that
// 2048 * ceil(n / 2^19), a multiple of 256 * 8 so that the kernels with 8 targets per thread apply; K <= 256 up to n = 2^27.
// On several GPUs the units of a rank should still fill >= 40 waves of its 2 x 148 resident CTAs, or the last partial
// wave shows (measured at n = 1 M on 8 GPUs with B = 4096: 12.7 waves, 40.3 ms instead of 38.8): B shrinks (in steps of
// 2048) until K(K+1)/2 >= 40 * 296 * world. NBX_SYM_BLOCK1024=1 restores the former rule 1024 * ceil(n / 2^18) (experiments).
uint32_t all_pairs_sym_block(uint32_t n, int world) {
  static const bool old_rule = [] { const char* v = getenv("NBX_SYM_BLOCK1024"); return v && atoi(v); }();
  if (n < (1u << 17) || old_rule) {
    const uint32_t mult = (n + (1u << 18) - 1) >> 18;
    return 1024u * (mult ? mult : 1);
  }
  uint32_t B = 2048u * ((n + (1u << 19) - 1) >> 19);
  if (world > 1) {
    const uint64_t want = uint64_t(40) * 296 * uint64_t(world);
    while (B > 2048u) {
      const uint64_t K = (uint64_t(n) + B - 1) / B;
      if (K * (K + 1) / 2 >= want) break;
      B -= 2048u;
    }
  }
  return B;
}

template <typename T, int D, int RI, int MINB, bool PACKED = false>
static int sym_launch(nbx_engine* e, bool fuse, int nc) {
  SymState* s = static_cast<SymState*>(e->sym);
  const uint32_t world = uint32_t(e->cfg.world_size), rank = uint32_t(e->cfg.rank);
  if (!s) {
    s       = new SymState();
    e->sym  = s;
    s->B    = all_pairs_sym_block(e->n, e->cfg.world_size);
    s->K    = (e->n + s->B - 1) / s->B;
    s->slab = uint64_t(s->K) * s->B;
    // partial sums of THIS rank's units only: 2 x B records per unit = about K * n / world records (3.9 GB at n = 1 M float
    // on one GPU). Checked against the free memory first, so a large n fails with a clear message instead of a bare
    // cudaMalloc error (the ordered kernel, NBX_FLAG_ALLPAIRS_ORDERED, needs no such buffer).
    const uint64_t units = uint64_t(s->K) * (s->K + 1) / 2;
    const uint64_t mine  = (units > rank ? units - rank + world - 1 : 0) / world;
    const size_t bytes   = sizeof(vec4_t<T>) * size_t(2) * s->B * size_t(mine ? mine : 1);
    size_t free_b = 0, total_b = 0;
    NBX_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // NBX_SYM_MAX_MB caps the buffer (tests of the fallback below)
    const char* cap_env = getenv("NBX_SYM_MAX_MB");  // read per allocation (once per engine)
    const size_t cap_mb = cap_env ? size_t(atoll(cap_env)) : size_t(0);
    if (bytes + (size_t(512) << 20) > free_b || (cap_mb && (bytes >> 20) >= cap_mb)) {
      delete s;
      e->sym = nullptr;
      return fail(NBX_ERR_CAPACITY, "symmetric all-pairs: the partial-sum buffer needs " + std::to_string(bytes >> 20) + " MiB, " +
                                        std::to_string(free_b >> 20) + " MiB are free; create the engine with NBX_FLAG_ALLPAIRS_ORDERED");
    }
    NBX_CUDA(cudaMalloc(&s->P, bytes));
    NBX_CUDA(cudaMalloc(&s->asum, sizeof(vec4_t<T>) * e->n_pad));
  }
  if constexpr (PACKED) {
    // the tiles of block J cover [J*B, (J+1)*B) <= K*B records; the position buffers are padded to max(n_pad, K*B) + 1024
    const uint64_t count = s->slab;
    if (!s->soa) {
      s->soa_stride = count;
      NBX_CUDA(cudaMalloc(&s->soa, sizeof(float) * 4 * count));
    }
    aos_to_soa_kernel<<<unsigned((count + 255) / 256), 256, 0, e->stream>>>(static_cast<const float4*>(e->xm[e->cur]), count, s->soa_stride,
                                                                            s->soa);
    e->launches++;
  }
  auto kern = [] {
    if constexpr (PACKED) return all_pairs_sym_packed_kernel<D, RI, MINB>;
    else return all_pairs_sym_kernel<T, D, RI, MINB>;
  }();
  const size_t smem = size_t(SYM_STAGES) * SYM_JT * sizeof(vec4_t<T>) + size_t(SYM_WARPS) * SYM_JT * 3 * sizeof(T) + SYM_STAGES * sizeof(uint64_t);
  NBX_TRY(ensure_dynamic_smem(e, kern, smem));
  const uint64_t units = uint64_t(s->K) * (s->K + 1) / 2;
  const uint32_t mine  = uint32_t((units > rank ? units - rank + world - 1 : 0) / world);
  SymArgs<T> p;
  p.xm = static_cast<const vec4_t<T>*>(e->xm[e->cur]);
  p.soa        = s->soa;
  p.soa_stride = s->soa_stride;
  p.P  = static_cast<vec4_t<T>*>(s->P);
  p.B = s->B;
  p.K = s->K;
  p.unit_begin  = rank;
  p.unit_stride = world;
  if (mine) kern<<<mine, 256, smem, e->stream>>>(p);
  e->launches++;
  LeapArgs<T> leap;
  leap.xm_in  = static_cast<const vec4_t<T>*>(e->xm[e->cur]);
  leap.xm_out = static_cast<vec4_t<T>*>(e->xm[e->cur ^ 1]);
  leap.v  = static_cast<vec4_t<T>*>(e->v);
  leap.a  = static_cast<vec4_t<T>*>(e->a);
  leap.ao = static_cast<vec4_t<T>*>(e->ao);
  leap.dt = T(e->cfg.dt);
  const unsigned gb = (e->n + 255) / 256;
  if (world == 1) {
    sym_reduce_kernel<T, D><<<gb, 256, 0, e->stream>>>(p.P, s->B, s->K, e->n, 0, 1, T(e->cfg.G), 1, fuse ? 1 : 0, nc,
                                                      static_cast<vec4_t<T>*>(s->asum), leap);
    e->launches++;
  } else {
    sym_reduce_kernel<T, D><<<gb, 256, 0, e->stream>>>(p.P, s->B, s->K, e->n, rank, world, T(e->cfg.G), 0, 0, nc,
                                                      static_cast<vec4_t<T>*>(s->asum), leap);
    NBX_TRY(comm_allreduce_sum(e, s->asum, size_t(e->n) * 4));
    sym_finish_kernel<T, D><<<gb, 256, 0, e->stream>>>(static_cast<const vec4_t<T>*>(s->asum), e->n, T(e->cfg.G), fuse ? 1 : 0, nc, leap);
    e->launches += 2;
  }
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

bool all_pairs_sym_enabled(const nbx_engine* e) {
  if (e->sym_unavailable) return false;
  if ((e->algo != NBX_ALL_PAIRS && e->algo != NBX_ALL_PAIRS_COLLAPSED) || (e->cfg.flags & NBX_FLAG_ALLPAIRS_ORDERED)) return false;
  if (e->cfg.flags & NBX_FLAG_ALLPAIRS_SYMMETRIC) return true;
  return e->n >= 16384;  // measured cross-over on B200; below that there are too few (I, J) units to fill 148 SMs
}

// collapsed_nc: 0 = all_pairs_force semantics; 2 / 3 = all_pairs_collapsed_force semantics over that many components
int all_pairs_sym_force(nbx_engine* e, bool fuse, int collapsed_nc) {
  if (e->prec == 4) {
    // FP32x2 kernel, 8 targets per thread where the block size allows it (B a multiple of 256*8), else 4. Measured at
    // n = 1 M: scalar 361 ms, packed x4 321 ms, packed x8 311 ms per step. NBX_SYM_PACKED=0|1|2 forces scalar / x4 / x8.
    const char* env  = getenv("NBX_SYM_PACKED");
    const int forced = env ? atoi(env) : -1;
    const bool can8 = all_pairs_sym_block(e->n, e->cfg.world_size) % 2048u == 0;
    const int packed = forced >= 0 ? (forced == 2 && !can8 ? 1 : forced) : (can8 ? 2 : 1);
    if (packed == 2) return e->dim == 2 ? sym_launch<float, 2, 8, 2, true>(e, fuse, collapsed_nc) : sym_launch<float, 3, 8, 2, true>(e, fuse, collapsed_nc);
    if (packed == 1) return e->dim == 2 ? sym_launch<float, 2, 4, 2, true>(e, fuse, collapsed_nc) : sym_launch<float, 3, 4, 2, true>(e, fuse, collapsed_nc);
    return e->dim == 2 ? sym_launch<float, 2, 4, 2>(e, fuse, collapsed_nc) : sym_launch<float, 3, 4, 2>(e, fuse, collapsed_nc);
  }
  // double: 4 targets per thread at one CTA per SM measured 7.6 % faster than 2 targets at two CTAs per SM (the
  // shuffle butterfly is amortised over twice as many pairs), 8 targets another 4 %
  {
    // 8 targets per thread where B allows it: the butterfly and the tile flush are amortised over twice as many pairs
    // (n = 262144: 59.4 -> 56.9 ms per step, FP64-pipe fraction 0.68 -> 0.71). NBX_SYM_DOUBLE_RI=4 forces 4 (experiments).
    const char* env = getenv("NBX_SYM_DOUBLE_RI");
    if (!(env && atoi(env) == 4) && all_pairs_sym_block(e->n, e->cfg.world_size) % 2048u == 0)
      return e->dim == 2 ? sym_launch<double, 2, 8, 1>(e, fuse, collapsed_nc) : sym_launch<double, 3, 8, 1>(e, fuse, collapsed_nc);
  }
  return e->dim == 2 ? sym_launch<double, 2, 4, 1>(e, fuse, collapsed_nc) : sym_launch<double, 3, 4, 1>(e, fuse, collapsed_nc);
}

void all_pairs_sym_destroy(nbx_engine* e) {
  SymState* s = static_cast<SymState*>(e->sym);
  if (!s) return;
  if (s->P) cudaFree(s->P);
  if (s->asum) cudaFree(s->asum);
  if (s->soa) cudaFree(s->soa);
  delete s;
  e->sym = nullptr;
}

}  // namespace nbx
