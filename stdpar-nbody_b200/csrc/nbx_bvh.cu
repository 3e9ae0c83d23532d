// nbx_bvh.cu — Hilbert-sorted implicit-BVH Barnes-Hut for sm_100a.
//
//   bbox kernels        : reference bounding_box            (src/bvh.h:17-22, src/vec.h:381-405)   [K10]
//   hilbert_keys_kernel : reference hilbert_sort, key part  (src/bvh.h:33-45, src/vec.h:266-356)   [K11]
//   sort_pairs          : reference std::sort               (src/bvh.h:62-69)  -> nbx_sort.cu     [K12]
//   gather_kernel       : reference permutation copy        (src/bvh.h:71-91)                      [K13]
//   build kernels       : reference bvh::build_tree         (src/bvh.h:175-244)                    [K14/K15]
//   force_kernel        : reference bvh::compute_force      (src/bvh.h:246-324)                    [K16]
//
// Everything that feeds an integer artefact or a tree node (bbox, cell index, key, centre of mass, AABB, width) is
// computed with explicit round-to-nearest intrinsics in the reference's association order (no FMA contraction), so
// keys, permutation and node arrays are BIT-EXACT against the pinned oracle. The traversal evaluates the acceptance
// test `bw^2 < theta^2 * dist2` with the same exact arithmetic (identical interaction sets) and only the accumulated
// force uses the fast MUFU path (tolerance-level difference).
#include <cfloat>
#include <cstdlib>

#include "nbx_internal.cuh"
#include "nbx_math.cuh"

namespace nbx {

namespace {

__device__ __forceinline__ uint32_t to_u32_sat(float q) { return __float2uint_rz(q); }    // cvt.rzi.u32.f32 saturates
__device__ __forceinline__ uint32_t to_u32_sat(double q) { return __double2uint_rz(q); }  // 

// FP32x2 ops with EXPLICIT .rn in PTX: each half is rounded exactly like the scalar __fmul_rn / __fadd_rn, and neither
// NVVM nor ptxas may contract them into an FFMA2 (the __fmul2_rn / __fadd2_rn intrinsics of sm_100_rt.h ARE contracted:
// checked in SASS) — required wherever a decision must be bit-identical to the reference's unfused arithmetic.
__device__ __forceinline__ float2 mul2_exact(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 add2_exact(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}

template <typename T>
__host__ __device__ constexpr T eps_of() {
  return sizeof(T) == 4 ? T(FLT_EPSILON) : T(DBL_EPSILON);
}
// vec.h:388: tol = (T)(epsilon * 10.)  (double product, then converted)
template <typename T>
__host__ __device__ inline T aabb_tol() {
  return sizeof(T) == 4 ? T(double(FLT_EPSILON) * 10.) : T(DBL_EPSILON * 10.);
}

template <typename T>
struct Box {  // device-resident result of the bbox reduction
  T lo[4], hi[4], cell[4];
};

// The walk's view of a node: centre of mass + mass + width^2 in one 32 B (float) / 64 B (double) aligned record, so a
// warp step costs one address computation and touches one sector-aligned line instead of two arrays.
template <typename T>
struct alignas(8 * sizeof(T)) WalkRec {
  T x, y, z, m, w2, pad[3];
};

template <typename T>
__device__ __forceinline__ WalkRec<T> make_rec(vec4_t<T> cm, T w2) {
  WalkRec<T> r;
  r.x = cm.x; r.y = cm.y; r.z = cm.z; r.m = cm.w; r.w2 = w2;
  r.pad[0] = r.pad[1] = r.pad[2] = T(0);
  return r;
}

template <typename T>
struct BvhState {
  uint32_t levels = 0;      // log2(bit_ceil(n)); tree levels 0..levels-1, bodies are level `levels`
  uint64_t nnodes = 0;      // 2^levels - 1
  Box<T>* box     = nullptr;
  T* partial      = nullptr;  // [blocks][8] partial min/max
  uint32_t nblocks = 0;
  uint64_t* keys   = nullptr;  // keys in entry order
  uint32_t* perm   = nullptr;
  vec4_t<T>* node_m = nullptr;  // (com.x, com.y, com.z, mass)
  T* bw             = nullptr;
  WalkRec<T>* rec   = nullptr;  // what the traversal loads: (com, mass, bw*bw) of node k in ONE aligned record; bw*bw is
                                // rounded once at build time, as bvh.h:247 rounds it per test
  vec4_t<T>* lo     = nullptr;
  vec4_t<T>* hi     = nullptr;
  bool have_box = false, sorted = false, built = false;
  int sort_mode = 0;  // multi-GPU: 0 sharded, 1 replicated, 2 broadcast (NBX_BVH_SORT, read when the engine is created)
};

// ---- K10 bounding box ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T tmin(T a, T b) { return fmin(a, b); }
template <typename T>
__device__ __forceinline__ T tmax(T a, T b) { return fmax(a, b); }

template <typename T, int D>
__global__ void __launch_bounds__(256) bbox_partial_kernel(const vec4_t<T>* __restrict__ xm, uint32_t n, T* partial) {
  T lo[3] = {T(0), T(0), T(0)}, hi[3] = {T(0), T(0), T(0)};  // identity: the origin (bvh.h:20)
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    vec4_t<T> b = xm[i];
    lo[0] = tmin(lo[0], b.x); hi[0] = tmax(hi[0], b.x);
    lo[1] = tmin(lo[1], b.y); hi[1] = tmax(hi[1], b.y);
    if (D == 3) { lo[2] = tmin(lo[2], b.z); hi[2] = tmax(hi[2], b.z); }
  }
  __shared__ T red[8][6];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo[k] = tmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
      hi[k] = tmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
    }
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 3; ++k) { red[threadIdx.x >> 5][k] = lo[k]; red[threadIdx.x >> 5][3 + k] = hi[k]; }
  __syncthreads();
  if (threadIdx.x < 6) {
    T v = red[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? tmin(v, red[w][threadIdx.x]) : tmax(v, red[w][threadIdx.x]);
    partial[blockIdx.x * 8 + threadIdx.x] = v;
  }
}

// final reduce + inflate + grid cell size (bvh.h:33-34). min_i(x_i - tol) == (min_i x_i) - tol because x - tol is
// monotone under round-to-nearest, so reducing first and inflating once is exact.
template <typename T, int D>
__global__ void bbox_final_kernel(const T* partial, uint32_t nblocks, Box<T>* box) {
  // one warp: lanes stride over the per-CTA partials, then a shuffle reduction (min/max are order-independent => exact)
  const int lane = threadIdx.x;
  T lo[3] = {T(0), T(0), T(0)}, hi[3] = {T(0), T(0), T(0)};
  for (uint32_t b = lane; b < nblocks; b += 32)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      lo[k] = tmin(lo[k], partial[b * 8 + k]);
      hi[k] = tmax(hi[k], partial[b * 8 + 3 + k]);
    }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo[k] = tmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
      hi[k] = tmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
    }
  if (lane >= 3) return;
  const int k = lane;  // one thread per axis
  const T tol = aabb_tol<T>();
  const T l   = tmin(sub_rn(T(0), tol), sub_rn(lo[k], tol));
  const T h   = tmax(add_rn(T(0), tol), add_rn(hi[k], tol));
  const T cells = D == 2 ? T(0xffffffffu) : T(0x1fffffu);
  box->lo[k]   = l;
  box->hi[k]   = h;
  box->cell[k] = div_rn(sub_rn(h, l), cells);
}

// ---- K11 Hilbert keys ------------------------------------------------------------------------------------------
// Skilling, "Programming the Hilbert curve" (AIP Conf. Proc. 707, 2004), transform over TWO axes with `bits` bits — in
// 3-D too: the reference's hilbert<3> sets n = 2 (vec.h:328) and interleaves axis 2 untransformed.
__device__ __forceinline__ void skilling2(uint32_t& x0, uint32_t& x1, int bits) {
  const uint32_t M = 1u << (bits - 1);
  for (uint32_t Q = M; Q > 1; Q >>= 1) {
    const uint32_t P = Q - 1;
    if (x0 & Q) x0 ^= P;  // axis 0: invert (the exchange branch is a no-op for axis 0 with itself)
    if (x1 & Q) x0 ^= P;  // axis 1: invert ...
    else {                // ... or exchange low bits of axis 0 and 1
      const uint32_t t = (x0 ^ x1) & P;
      x0 ^= t;
      x1 ^= t;
    }
  }
  x1 ^= x0;  // Gray encode
  // the reference's loop `for (Q = M; Q > 1; Q >>= 1) if (x1 & Q) t ^= Q - 1;` (vec.h:318-323) sets bit j of t to the
  // parity of the bits of x1 above j, up to bit `bits`-1 (a cell index that rounded up to 2^bits carries a higher bit, which
  // the loop never looks at): a suffix XOR in five doubling steps, shifted down by one
  uint32_t t = x1 & ((M << 1) - 1u);
  t ^= t >> 1;
  t ^= t >> 2;
  t ^= t >> 4;
  t ^= t >> 8;
  t ^= t >> 16;
  t >>= 1;
  x0 ^= t;
  x1 ^= t;
}
__device__ __forceinline__ uint64_t spread2(uint64_t x) {  // vec.h:268-275
  x = (x | x << 16) & 0xffff0000ffffull;
  x = (x | x << 8) & 0xff00ff00ff00ffull;
  x = (x | x << 4) & 0xf0f0f0f0f0f0f0full;
  x = (x | x << 2) & 0x3333333333333333ull;
  x = (x | x << 1) & 0x5555555555555555ull;
  return x;
}
__device__ __forceinline__ uint64_t spread3(uint64_t x) {  // vec.h:278-286
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

template <typename T, int D>
__global__ void __launch_bounds__(256) hilbert_keys_kernel(const vec4_t<T>* __restrict__ xm, uint32_t n,
                                                           const Box<T>* __restrict__ box, uint64_t* __restrict__ keys) {
  uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  vec4_t<T> b = xm[i];
  // bvh.h:42: cast<uint32_t>((x - mins) / grid_cell_size), IEEE subtract + divide, truncating (saturating) convert
  uint32_t c0 = to_u32_sat(div_rn(sub_rn(b.x, box->lo[0]), box->cell[0]));
  uint32_t c1 = to_u32_sat(div_rn(sub_rn(b.y, box->lo[1]), box->cell[1]));
  if (D == 2) {
    skilling2(c0, c1, 32);
    keys[i] = spread2(c1) | (spread2(c0) << 1);  // vec.h:277
  } else {
    uint32_t c2 = to_u32_sat(div_rn(sub_rn(b.z, box->lo[2]), box->cell[2]));
    skilling2(c0, c1, 21);
    keys[i] = spread3(c2) | (spread3(c1) << 1) | (spread3(c0) << 2);  // vec.h:288
  }
}

// ---- K13 apply the permutation to all five arrays (bvh.h:71-91) ---------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gather_kernel(const uint32_t* __restrict__ perm, uint32_t n,
                                                     const vec4_t<T>* __restrict__ xm_in, vec4_t<T>* __restrict__ xm_out,
                                                     const vec4_t<T>* __restrict__ v_in, vec4_t<T>* __restrict__ v_out,
                                                     const vec4_t<T>* __restrict__ a_in, vec4_t<T>* __restrict__ a_out,
                                                     const vec4_t<T>* __restrict__ ao_in, vec4_t<T>* __restrict__ ao_out) {
  uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  uint32_t src = perm[i];
  xm_out[i] = xm_in[src];
  v_out[i]  = v_in[src];
  a_out[i]  = a_in[src];
  ao_out[i] = ao_in[src];
}

// ---- K14/K15 build ----------------------------------------------------------------------------------------------
template <typename T, int D>
__device__ __forceinline__ T node_width(vec4_t<T> lo, vec4_t<T> hi) {  // bvh.h:140-144
  T l0 = sub_rn(hi.x, lo.x), l1 = sub_rn(hi.y, lo.y);
  T w  = l0 < l1 ? l1 : l0;
  if (D == 3) {
    T l2 = sub_rn(hi.z, lo.z);
    w    = w < l2 ? l2 : w;
  }
  return w;
}

// deepest tree level from pairs of bodies (bvh.h:178-207)
template <typename T, int D>
__global__ void __launch_bounds__(256) build_leaf_level_kernel(const vec4_t<T>* __restrict__ xm, uint32_t n, uint32_t first,
                                                               uint32_t count, vec4_t<T>* node_m, T* bw, WalkRec<T>* rec, vec4_t<T>* lo,
                                                               vec4_t<T>* hi) {
  uint32_t li = blockIdx.x * 256 + threadIdx.x;
  if (li >= count) return;
  const uint32_t i = first + li, bl = li * 2, br = bl + 1;
  const T tol = aabb_tol<T>();
  if (bl >= n) {
    node_m[i] = make_v4<T>(0, 0, 0, 0);  // dead node (mass 0); b/bw stay as allocated (zero)
    rec[i]    = make_rec<T>(make_v4<T>(0, 0, 0, 0), T(0));
    return;
  }
  vec4_t<T> pl = xm[bl];
  if (br >= n) {
    node_m[i]     = pl;
    vec4_t<T> blo = make_v4<T>(sub_rn(pl.x, tol), sub_rn(pl.y, tol), D == 3 ? sub_rn(pl.z, tol) : T(0), 0);
    vec4_t<T> bhi = make_v4<T>(add_rn(pl.x, tol), add_rn(pl.y, tol), D == 3 ? add_rn(pl.z, tol) : T(0), 0);
    lo[i] = blo; hi[i] = bhi;
    const T w = node_width<T, D>(blo, bhi);
    bw[i] = w; rec[i] = make_rec<T>(pl, mul_rn(w, w));
    return;
  }
  vec4_t<T> pr = xm[br];
  T mass       = add_rn(pl.w, pr.w);
  vec4_t<T> c;
  c.x = div_rn(add_rn(mul_rn(pl.w, pl.x), mul_rn(pr.w, pr.x)), mass);
  c.y = div_rn(add_rn(mul_rn(pl.w, pl.y), mul_rn(pr.w, pr.y)), mass);
  c.z = D == 3 ? div_rn(add_rn(mul_rn(pl.w, pl.z), mul_rn(pr.w, pr.z)), mass) : T(0);
  c.w = mass;
  node_m[i]     = c;
  vec4_t<T> blo = make_v4<T>(sub_rn(tmin(pl.x, pr.x), tol), sub_rn(tmin(pl.y, pr.y), tol),
                             D == 3 ? sub_rn(tmin(pl.z, pr.z), tol) : T(0), 0);
  vec4_t<T> bhi = make_v4<T>(add_rn(tmax(pl.x, pr.x), tol), add_rn(tmax(pl.y, pr.y), tol),
                             D == 3 ? add_rn(tmax(pl.z, pr.z), tol) : T(0), 0);
  lo[i] = blo; hi[i] = bhi;
  const T w = node_width<T, D>(blo, bhi);
  bw[i] = w; rec[i] = make_rec<T>(c, mul_rn(w, w));
}

// one node of an upper level from its children (bvh.h:210-243)
template <typename T, int D>
__device__ __forceinline__ void build_node(uint32_t first, uint32_t count, uint32_t li, vec4_t<T>* node_m, T* bw, WalkRec<T>* rec,
                                           vec4_t<T>* lo, vec4_t<T>* hi) {
  const uint32_t i = first + li, bl = li * 2 + first + count, br = bl + 1;
  vec4_t<T> ml = node_m[bl], mr = node_m[br];
  if (!(ml.w != T(0))) {
    node_m[i] = ml;
    rec[i]    = make_rec<T>(ml, T(0));  // the reference leaves b/bw of dead nodes as allocated (zero)
    return;
  }
  if (!(mr.w != T(0))) {
    node_m[i] = ml;
    lo[i] = lo[bl]; hi[i] = hi[bl];
    bw[i] = bw[bl]; rec[i] = rec[bl];
    return;
  }
  T mass = add_rn(ml.w, mr.w);
  vec4_t<T> c;
  c.x = div_rn(add_rn(mul_rn(ml.w, ml.x), mul_rn(mr.w, mr.x)), mass);
  c.y = div_rn(add_rn(mul_rn(ml.w, ml.y), mul_rn(mr.w, mr.y)), mass);
  c.z = D == 3 ? div_rn(add_rn(mul_rn(ml.w, ml.z), mul_rn(mr.w, mr.z)), mass) : T(0);
  c.w = mass;
  node_m[i] = c;
  vec4_t<T> l0 = lo[bl], l1 = lo[br], h0 = hi[bl], h1 = hi[br];
  vec4_t<T> blo = make_v4<T>(tmin(l0.x, l1.x), tmin(l0.y, l1.y), D == 3 ? tmin(l0.z, l1.z) : T(0), 0);
  vec4_t<T> bhi = make_v4<T>(tmax(h0.x, h1.x), tmax(h0.y, h1.y), D == 3 ? tmax(h0.z, h1.z) : T(0), 0);
  lo[i] = blo; hi[i] = bhi;
  const T w = node_width<T, D>(blo, bhi);
  bw[i] = w; rec[i] = make_rec<T>(c, mul_rn(w, w));
}

template <typename T, int D>
__global__ void __launch_bounds__(256) build_level_kernel(uint32_t first, uint32_t count, vec4_t<T>* node_m, T* bw, WalkRec<T>* rec,
                                                          vec4_t<T>* lo, vec4_t<T>* hi) {
  const uint32_t li = blockIdx.x * 256 + threadIdx.x;
  if (li < count) build_node<T, D>(first, count, li, node_m, bw, rec, lo, hi);
}

// the top of the tree, levels `top` .. 0 (at most 1024 nodes each), in ONE launch: a single CTA walks up level by level
// (the reference launches one parallel algorithm per level, bvh.h:210-243)
constexpr uint32_t BVH_TOP_LEVEL = 10;
template <typename T, int D>
__global__ void __launch_bounds__(1024) build_top_levels_kernel(int top, vec4_t<T>* node_m, T* bw, WalkRec<T>* rec, vec4_t<T>* lo,
                                                                vec4_t<T>* hi) {
  for (int l = top; l >= 0; --l) {
    const uint32_t count = 1u << l;
    if (threadIdx.x < count) build_node<T, D>(count - 1, count, threadIdx.x, node_m, bw, rec, lo, hi);
    __syncthreads();  // the next level reads what this one wrote (same CTA: visible after the barrier)
  }
}

// ---- K16 traversal ------------------------------------------------------------------------------------------------
// dist2 in the reference's order: ((dx*dx) + dy*dy) + dz*dz with d = xs - xj, no contraction (vec.h:232-240)
template <typename T, int D>
__device__ __forceinline__ T dist2_exact(T ax, T ay, T az, T bx, T by, T bz) {
  T dx = sub_rn(ax, bx), dy = sub_rn(ay, by);
  T r  = add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
  if (D == 3) {
    T dz = sub_rn(az, bz);
    r    = add_rn(r, mul_rn(dz, dz));
  }
  return r;
}

// One thread per (Hilbert-sorted) body, the reference's stackless walk over the implicit complete binary tree in heap
// order (root 0, children 2k+1 / 2k+2): "ascend right" of a left child k is k+1, of a right child k is k/2 one level up
// (= parent+1, bvh.h:272-281), leaving the body level goes to (k+1)/2. Neighbouring lanes hold neighbouring bodies of
// the Hilbert order and follow nearly the same path, so node loads are mostly warp-uniform broadcasts.
template <typename T, int D>
__global__ void __launch_bounds__(128) bvh_force_kernel(const vec4_t<T>* __restrict__ xm, const vec4_t<T>* __restrict__ node_m,
                                                        const T* __restrict__ bw, uint32_t n, uint32_t tb, uint32_t te,
                                                        uint32_t levels, T theta2, T c, vec4_t<T>* __restrict__ a_out) {
  const uint32_t i = tb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= te) return;
  const vec4_t<T> xs     = xm[i];
  const uint32_t leaf_lv = levels;  // bodies
  const uint32_t leaf0   = (1u << leaf_lv) - 1;
  uint32_t k = 0, level = 0;
  uint64_t covered = 0;
  T ax = 0, ay = 0, az = 0;
  while (covered < n) {
    if (level == leaf_lv) {
      uint32_t bidx = k - leaf0;
#pragma unroll
      for (int q = 0; q < 2; ++q, ++bidx) {
        if (bidx < n && bidx != i) {
          vec4_t<T> b = xm[bidx];
          T d2 = dist2_exact<T, D>(xs.x, xs.y, xs.z, b.x, b.y, b.z);
          T s  = b.w * inv_dist3(d2);
          ax = fma(b.x - xs.x, s, ax);
          ay = fma(b.y - xs.y, s, ay);
          if (D == 3) az = fma(b.z - xs.z, s, az);
        }
      }
      covered += 2;
      k = (k + 1) >> 1;
      level -= 1;
    } else {
      const vec4_t<T> nm = node_m[k];
      const T w          = bw[k];
      const T d2         = dist2_exact<T, D>(xs.x, xs.y, xs.z, nm.x, nm.y, nm.z);
      if (mul_rn(w, w) < mul_rn(theta2, d2)) {  // can_approximate (bvh.h:246-248)
        T s = nm.w * inv_dist3(d2);
        ax = fma(nm.x - xs.x, s, ax);
        ay = fma(nm.y - xs.y, s, ay);
        if (D == 3) az = fma(nm.z - xs.z, s, az);
        covered += uint64_t(1) << (levels - level);
        if (k & 1) k += 1;            // left child -> right sibling
        else { k >>= 1; level -= 1; }  // right child (or root) -> parent + 1, one level up
      } else {
        k = 2 * k + 1;
        level += 1;
      }
    }
  }
  a_out[i] = make_v4<T>(mul_rn(c, ax), mul_rn(c, ay), D == 3 ? mul_rn(c, az) : T(0), T(0));
}

// WARP-COOPERATIVE version of the same walk (used when the leaf offsets fit 27 bits): each lane keeps the reference's
// per-body state machine — its progress `covered` (= num_covered_particles, bvh.h:267) and the level of the node it wants
// next; the node index follows from both: k = 2^level - 1 + (covered >> (levels - level)). Every step the warp takes the
// smallest (covered, level) key over its bodies with one REDUX.MIN, loads THAT node once (warp-uniform address), and only
// the bodies that asked for it act on it. The warp therefore walks the union of its bodies' paths in leaf order; each body
// performs exactly the reference's sequence of tests and interactions.
//
// The per-body state is folded into ONE word. `covered` is always even (nodes cover >= 2 leaves, the
// body level covers 2), so key = (covered << 4) | level holds covered <= 2^27 without overflow. The key only ever grows:
// opening node (covered, level) goes to (covered, level + 1) = key + 1 — which no other body can be below, so the update
// is key = max(key, candidate) with a warp-uniform candidate and needs no "is this my node" masking — and accepting adds
// 2^(levels-level) leaves minus one level for a right child (sibling for a left child). The acceptance test is evaluated
// for every body (the node loads are warp-uniform anyway); only the accumulation is predicated. A body is finished when
// covered >= n, i.e. key >= nlim; the warp stops when the minimum is.
//
// NB bodies per lane (32*NB consecutive Hilbert-sorted bodies per warp; fewer lanes carry bodies at small n, see walk_lanes): the bookkeeping of a step — REDUX, node index,
// address, record load, candidates, loop control, about two thirds of the instructions — is shared by NB tests, and the
// union of 64 (128) neighbouring paths is only 1.12x (1.30x) longer than that of 32 (tools/bvh_walk_sim.c, n = 10 M:
// 10257 / 11499 / 13339 steps per warp for NB = 1 / 2 / 4). In float the NB tests run two at a time on FP32x2
// instructions (add/mul.rn.f32x2 round each half exactly like the scalar ops, and are never contracted), so decisions
// stay bit-identical to the reference's.
// the other ranks' acceleration arrays (CUDA IPC views, nbx_peer_import): the walk stores each finished body's result
// into every one of them — P2P stores over NVLink issued warp by warp while the other warps still walk — instead of an
// all-gather after the kernel. count = 0: single GPU, or the NCCL path.
struct PeerOut {
  void* p[15];
  int count;
};

// Small problems: pull the 32 B sector(s) of a left node's RIGHT SIBLING towards L1 together with the node's own load. Every
// body that visits the left child visits the right one as soon as the left subtree is done, and at small n the walk is a
// chain of dependent L2 round trips (about 300 cycles per step of the longest warp), not an issue-bound stream. Measured
// (B200, step ms without / with): float n = 10 k 0.459 / 0.442, 30 k 0.910 / 0.806, 100 k 1.761 / 1.536, 200 k 2.256 / 2.188,
// 400 k 3.93 / 4.02, 1 M 9.25 / 10.63; double 30 k 1.311 / 1.081, 100 k 2.481 / 2.089 — on below 250 k targets. Touching
// the left child as well (the next node whenever a body opens) LOSES 3-20 % at every size, and so did `prefetch.global.L1`.
// An ordinary load is used; ptxas deletes a load nobody reads (asm volatile or not), so the value is folded into a word
// that is consumed one step later, when it has long landed.
__device__ __forceinline__ unsigned touch_sector(const void* p) {
  unsigned v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <typename T, int D, int NB, bool COUNT = false, bool TOUCH = false>
__global__ void __launch_bounds__(128) bvh_force_key_kernel(const vec4_t<T>* __restrict__ xm, const WalkRec<T>* __restrict__ rec1,
                                                            uint32_t n, uint32_t tb, uint32_t te, uint32_t levels, T theta2, T c,
                                                            vec4_t<T>* __restrict__ a_out, unsigned long long* stats, const PeerOut peers,
                                                            const uint32_t lanes) {
  // rec1 is the record array offset by -1 element: indexed by the 1-based heap index kk = k + 1
  // lanes: bodies per warp and slot (32, or 16 / 8 for small problems: only the first `lanes` lanes carry bodies — the walk
  // ends with its heaviest warp, and the union path of 8 neighbours is a third shorter than that of 32)
  constexpr bool PACK = sizeof(T) == 4 && NB % 2 == 0;  // FP32x2 over pairs of bodies
  const uint32_t lane  = threadIdx.x & 31u;
  const uint32_t wbase = tb + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (lanes * NB);  // first body of the warp
  unsigned long long n_visit = 0, n_take = 0, n_step = 0;  // COUNT only
  // covered is even: covered >= n <=> covered >= n rounded up to even. Active keys are <= ((n_even - 2) << 4) + 27, so the
  // limit can sit 4 below n_even << 4: accepting the ROOT (a "right child" by parity of kk = 1) then needs no special case,
  // its candidate (2^levels << 4) - 1 is past the limit for every n <= 2^levels.
  const uint32_t nlim  = ((n + (n & 1u)) << 4) - 4u;
  const uint32_t sent  = 16u << levels;  // one above the largest active covered << 4
  const uint32_t step0 = 16u << levels;  // key increment of accepting a level-0 node; >> level for deeper ones
  uint32_t idx[NB], key[NB];
  T px[NB], py[NB], pz[NB], ax[NB], ay[NB], az[NB];
  unsigned t_pend = 0, t_sink = 0;  // TOUCH only
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    idx[j]            = wbase + j * lanes + lane;
    const bool valid  = lane < lanes && idx[j] < te;
    const vec4_t<T> b = xm[valid ? idx[j] : tb];
    px[j] = b.x; py[j] = b.y; pz[j] = b.z;
    ax[j] = ay[j] = az[j] = T(0);
    key[j] = valid ? 0u : 0xffffffffu;
  }
  for (;;) {
    uint32_t kmine = key[0];
#pragma unroll
    for (int j = 1; j < NB; ++j) kmine = min(kmine, key[j]);
    const uint32_t kmin = __reduce_min_sync(0xffffffffu, kmine);
    if (kmin >= nlim) break;
    const uint32_t cl = kmin & 31u;
    if (COUNT) n_step += 1;
    if (cl == levels) {  // body level: the two bodies cpos, cpos+1 (bvh.h:288-303)
      const uint32_t cpos = (kmin >> 4) & ~1u;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t bidx = cpos + q;
        if (bidx < n) {
          const vec4_t<T> b = xm[bidx];
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            if (key[j] == kmin && bidx != idx[j]) {
              const T dx = sub_rn(b.x, px[j]), dy = sub_rn(b.y, py[j]), dz = D == 3 ? sub_rn(b.z, pz[j]) : T(0);
              T d2 = add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
              if (D == 3) d2 = add_rn(d2, mul_rn(dz, dz));
              const T s = b.w * inv_dist3(d2);
              ax[j] = fma(dx, s, ax[j]);
              ay[j] = fma(dy, s, ay[j]);
              if (D == 3) az[j] = fma(dz, s, az[j]);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const bool act = key[j] == kmin;
        if (COUNT) { n_visit += act; n_take += act; }
        if (act) key[j] = kmin + (cl ? 31u : 32u);  // covered += 2, level -= 1 (n = 1: the body level is level 0)
      }
      continue;
    }
    const uint32_t kk = (kmin | sent) >> (levels + 4u - cl);  // 1-based heap index: 2^level + covered / 2^(levels-level)
    const WalkRec<T>* r = rec1 + kk;
    const vec4_t<T> nm  = *reinterpret_cast<const vec4_t<T>*>(r);  // (com, mass)
    const T w2          = r->w2;
    if constexpr (TOUCH) {
      t_sink ^= t_pend;  // last step's touch: landed long ago, no wait
      if (!(kk & 1u)) {
        t_pend = touch_sector(r + 1);
        if (sizeof(T) == 8) t_pend ^= touch_sector(&r[1].w2);
      }
    }
    // accept: covered += 2^(levels-level); a right child (kk odd) continues one level up, a left child with its sibling
    const uint32_t cand_take = kmin + (step0 >> cl) - (kk & 1u);
    const uint32_t cand_open = kmin + 1u;
    if constexpr (PACK) {
      const float2 nx2 = make_float2(nm.x, nm.x), ny2 = make_float2(nm.y, nm.y), nz2 = make_float2(nm.z, nm.z);
      const float2 th2 = make_float2(theta2, theta2);
#pragma unroll
      for (int j = 0; j < NB; j += 2) {
        // (xj - xs) == -(xs - xj) exactly, so one difference serves the reference-order dist2 and the accumulation
        const float2 dx = add2_exact(nx2, make_float2(-px[j], -px[j + 1]));
        const float2 dy = add2_exact(ny2, make_float2(-py[j], -py[j + 1]));
        float2 dz       = make_float2(0.f, 0.f);
        // the squares are packed, their sums scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with
        // explicit .rn (checked in SASS), which would change the rounding of dist2; scalar add.rn is never contracted
        const float2 sx = mul2_exact(dx, dx), sy = mul2_exact(dy, dy);
        float2 d2       = make_float2(__fadd_rn(sx.x, sy.x), __fadd_rn(sx.y, sy.y));
        if (D == 3) {
          dz              = add2_exact(nz2, make_float2(-pz[j], -pz[j + 1]));
          const float2 sz = mul2_exact(dz, dz);
          d2              = make_float2(__fadd_rn(d2.x, sz.x), __fadd_rn(d2.y, sz.y));
        }
        const float2 t   = mul2_exact(th2, d2);
        const bool take0 = (key[j] == kmin) & (w2 < t.x);      // can_approximate (bvh.h:246-248)
        const bool take1 = (key[j + 1] == kmin) & (w2 < t.y);
        if (COUNT) { n_visit += (key[j] == kmin) + (key[j + 1] == kmin); n_take += take0 + take1; }
        if (take0 | take1) {
          float2 sq, inv;
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq.x) : "f"(d2.x));
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq.y) : "f"(d2.y));
          const float2 den = __ffma2_rn(d2, sq, make_float2(FLT_EPSILON, FLT_EPSILON));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(den.x));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(den.y));
          const float2 s = __fmul2_rn(make_float2(take0 ? nm.w : 0.f, take1 ? nm.w : 0.f), inv);
          float2 acc;
          acc = __ffma2_rn(dx, s, make_float2(ax[j], ax[j + 1])); ax[j] = acc.x; ax[j + 1] = acc.y;
          acc = __ffma2_rn(dy, s, make_float2(ay[j], ay[j + 1])); ay[j] = acc.x; ay[j + 1] = acc.y;
          if (D == 3) { acc = __ffma2_rn(dz, s, make_float2(az[j], az[j + 1])); az[j] = acc.x; az[j + 1] = acc.y; }
        }
        key[j]     = max(key[j], take0 ? cand_take : cand_open);
        key[j + 1] = max(key[j + 1], take1 ? cand_take : cand_open);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const T dx = sub_rn(nm.x, px[j]), dy = sub_rn(nm.y, py[j]), dz = D == 3 ? sub_rn(nm.z, pz[j]) : T(0);
        T d2 = add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
        if (D == 3) d2 = add_rn(d2, mul_rn(dz, dz));
        const bool act  = key[j] == kmin;
        const bool take = act & (w2 < mul_rn(theta2, d2));  // can_approximate (bvh.h:246-248)
        if (COUNT) { n_visit += act; n_take += take; }
        if (take) {
          const T s = nm.w * inv_dist3(d2);
          ax[j] = fma(dx, s, ax[j]);
          ay[j] = fma(dy, s, ay[j]);
          if (D == 3) az[j] = fma(dz, s, az[j]);
        }
        key[j] = max(key[j], take ? cand_take : cand_open);
      }
    }
  }
  if constexpr (COUNT) {
    atomicAdd(&stats[0], n_visit);
    atomicAdd(&stats[1], n_take);
    if (lane == 0) atomicAdd(&stats[2], n_step);
  } else {
#pragma unroll
    for (int j = 0; j < NB; ++j)
      if (lane < lanes && idx[j] < te) {
        const vec4_t<T> r = make_v4<T>(mul_rn(c, ax[j]), mul_rn(c, ay[j]), D == 3 ? mul_rn(c, az[j]) : T(0), T(0));
        a_out[idx[j]] = r;
        for (int q = 0; q < peers.count; ++q) static_cast<vec4_t<T>*>(peers.p[q])[idx[j]] = r;
      }
    if (peers.count) __threadfence_system();
    if constexpr (TOUCH)
      if (t_sink == 0x7fc5a5a5u && n == 0u) a_out[0].x = T(0);  // never true (n > 0); keeps the touches alive
  }
}

// ---- artefact export ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void export_nodes_kernel(const vec4_t<T>* node_m, const T* bw, const vec4_t<T>* lo, const vec4_t<T>* hi,
                                    uint64_t first, uint64_t count, T* out_m, T* out_bw, T* out_b) {
  uint64_t q = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (q >= count) return;
  uint64_t i     = first + q;
  vec4_t<T> m    = node_m[i];
  T* om          = out_m + q * (D + 1);
  om[0] = m.x; om[1] = m.y;
  if (D == 3) om[2] = m.z;
  om[D]     = m.w;
  out_bw[q] = bw[i];
  vec4_t<T> l = lo[i], h = hi[i];
  T* ob = out_b + q * 2 * D;
  ob[0] = l.x; ob[1] = l.y;
  if (D == 3) ob[2] = l.z;
  ob[D] = h.x; ob[D + 1] = h.y;
  if (D == 3) ob[D + 2] = h.z;
}

template <typename T>
BvhState<T>* st(nbx_engine* e) { return static_cast<BvhState<T>*>(e->bvh); }

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------------
template <typename T, int D>
static int create_impl(nbx_engine* e) {
  auto* s = new BvhState<T>();
  e->bvh  = s;
  uint32_t nleafs = 1, levels = 0;
  while (nleafs < e->n) { nleafs <<= 1; ++levels; }
  if (levels > 30) return fail(NBX_ERR_INVALID, "bvh: n too large");
  s->levels  = levels;
  {
    const char* v = getenv("NBX_BVH_SORT");
    const std::string m = v ? v : "sharded";
    s->sort_mode = m == "broadcast" ? 2 : (m == "replicated" ? 1 : 0);
  }
  s->nnodes  = (uint64_t(1) << levels) - 1;
  s->nblocks = std::min<uint32_t>((e->n + 255) / 256, uint32_t(e->sm_count) * 8);
  NBX_CUDA(cudaMalloc(&s->box, sizeof(Box<T>)));
  NBX_CUDA(cudaMalloc(&s->partial, sizeof(T) * 8 * s->nblocks));
  NBX_CUDA(cudaMalloc(&s->keys, sizeof(uint64_t) * size_t(e->n)));
  NBX_CUDA(cudaMalloc(&s->perm, sizeof(uint32_t) * size_t(e->n)));
  const size_t nn = size_t(s->nnodes ? s->nnodes : 1);
  NBX_CUDA(cudaMalloc(&s->node_m, sizeof(vec4_t<T>) * nn));
  NBX_CUDA(cudaMalloc(&s->bw, sizeof(T) * nn));
  NBX_CUDA(cudaMalloc(&s->rec, sizeof(WalkRec<T>) * nn));
  NBX_CUDA(cudaMalloc(&s->lo, sizeof(vec4_t<T>) * nn));
  NBX_CUDA(cudaMalloc(&s->hi, sizeof(vec4_t<T>) * nn));
  // the reference never writes b/bw of dead nodes (bvh.h:185-188,225-228): keep them deterministic (zero)
  NBX_CUDA(cudaMemsetAsync(s->node_m, 0, sizeof(vec4_t<T>) * nn, e->stream));
  NBX_CUDA(cudaMemsetAsync(s->bw, 0, sizeof(T) * nn, e->stream));
  NBX_CUDA(cudaMemsetAsync(s->rec, 0, sizeof(WalkRec<T>) * nn, e->stream));
  NBX_CUDA(cudaMemsetAsync(s->lo, 0, sizeof(vec4_t<T>) * nn, e->stream));
  NBX_CUDA(cudaMemsetAsync(s->hi, 0, sizeof(vec4_t<T>) * nn, e->stream));
  const size_t rb = rec_bytes(e);
  void** alts[3]  = {&e->v_alt, &e->a_alt, &e->ao_alt};
  for (auto pp : alts) {
    NBX_CUDA(cudaMalloc(pp, rb * e->n_pad));
    NBX_CUDA(cudaMemsetAsync(*pp, 0, rb * e->n_pad, e->stream));
  }
  e->own_a[0] = e->a;
  e->own_a[1] = e->a_alt;
  NBX_TRY(sorter_create(e, e->n));
  return NBX_OK;
}

template <typename T, int D>
static void destroy_impl(nbx_engine* e) {
  auto* s = st<T>(e);
  if (!s) return;
  void* bufs[] = {s->box, s->partial, s->keys, s->perm, s->node_m, s->bw, s->rec, s->lo, s->hi};
  for (void* b : bufs)
    if (b) cudaFree(b);
  delete s;
  e->bvh = nullptr;
}

template <typename T, int D>
static int bbox_impl(nbx_engine* e) {
  auto* s = st<T>(e);
  bbox_partial_kernel<T, D><<<s->nblocks, 256, 0, e->stream>>>(static_cast<const vec4_t<T>*>(e->xm[e->cur]), e->n, s->partial);
  bbox_final_kernel<T, D><<<1, 32, 0, e->stream>>>(s->partial, s->nblocks, s->box);
  e->launches += 2;
  NBX_CUDA(cudaGetLastError());
  s->have_box = true;
  return NBX_OK;
}

template <typename T, int D>
static int sort_impl(nbx_engine* e) {
  auto* s = st<T>(e);
  if (!s->have_box) return fail(NBX_ERR_STATE, "hilbert_sort before bounding_box");
  const uint32_t n = e->n;
  // Multi-GPU (measured in DESIGN.md §7): NBX_BVH_SORT = sharded (default: the ranks split the key range,
  // sort_pairs_sharded) | replicated (every rank sorts all n keys; deterministic => identical permutations) | broadcast
  // (rank 0 sorts, the permutation is broadcast over NVLink: the other ranks only wait). All three give the same bits.
  const int mode = s->sort_mode;
  const bool multi = e->cfg.world_size > 1;
  const int bits   = D == 2 ? 64 : 63;
  if (!(multi && mode == 2 && e->cfg.rank != 0)) {
    hilbert_keys_kernel<T, D><<<(n + 255) / 256, 256, 0, e->stream>>>(static_cast<const vec4_t<T>*>(e->xm[e->cur]), n, s->box, s->keys);
    e->launches++;
    if (multi && mode == 0) NBX_TRY(sort_pairs_sharded(e, s->keys, n, bits, s->perm));
    else NBX_TRY(sort_pairs(e, s->keys, n, bits, s->perm, nullptr));
  }
  if (multi && mode == 2) NBX_TRY(comm_broadcast(e, s->perm, sizeof(uint32_t) * size_t(n), 0));
  gather_kernel<T><<<(n + 255) / 256, 256, 0, e->stream>>>(
      s->perm, n, static_cast<const vec4_t<T>*>(e->xm[e->cur]), static_cast<vec4_t<T>*>(e->xm[e->cur ^ 1]),
      static_cast<const vec4_t<T>*>(e->v), static_cast<vec4_t<T>*>(e->v_alt), static_cast<const vec4_t<T>*>(e->a),
      static_cast<vec4_t<T>*>(e->a_alt), static_cast<const vec4_t<T>*>(e->ao), static_cast<vec4_t<T>*>(e->ao_alt));
  e->launches++;
  NBX_CUDA(cudaGetLastError());
  e->cur ^= 1;
  std::swap(e->v, e->v_alt);
  std::swap(e->a, e->a_alt);
  std::swap(e->ao, e->ao_alt);
  s->sorted = true;
  return NBX_OK;
}

template <typename T, int D>
static int build_impl(nbx_engine* e) {
  auto* s = st<T>(e);
  if (s->levels == 0) {  // n = 1: no tree nodes, the walk only meets the body level (bvh.h:288-303)
    s->built = true;
    return NBX_OK;
  }
  const vec4_t<T>* xm = static_cast<const vec4_t<T>*>(e->xm[e->cur]);
  const uint32_t last = s->levels - 1;
  {
    uint32_t first = (1u << last) - 1, count = 1u << last;
    build_leaf_level_kernel<T, D><<<(count + 255) / 256, 256, 0, e->stream>>>(xm, e->n, first, count, s->node_m, s->bw, s->rec, s->lo, s->hi);
    e->launches++;
  }
  int l = int(last) - 1;
  for (; l > int(BVH_TOP_LEVEL); --l) {
    uint32_t first = (1u << l) - 1, count = 1u << l;
    build_level_kernel<T, D><<<(count + 255) / 256, 256, 0, e->stream>>>(first, count, s->node_m, s->bw, s->rec, s->lo, s->hi);
    e->launches++;
  }
  if (l >= 0) {
    build_top_levels_kernel<T, D><<<1, 1024, 0, e->stream>>>(l, s->node_m, s->bw, s->rec, s->lo, s->hi);
    e->launches++;
  }
  NBX_CUDA(cudaGetLastError());
  s->built = true;
  return NBX_OK;
}

// bodies per lane of the warp walk. More bodies per lane share the bookkeeping of a step but leave fewer warps; small
// problems are bound by the per-warp latency of the walk and want the most warps. Measured walk times (ms, B200, 3-D
// galaxy) for 1 / 2 / 4 bodies per lane — float: n = 100 k 1.71 / 2.35 / 3.16, 1 M 9.52 / 9.88 / 9.91, 3 M 33.9 / 30.2 /
// 28.7, 10 M 137.0 / 118.2 / 110.2; double: 100 k 2.66 / 3.22 / 4.56, 1 M 15.6 / 17.0 / 18.8, 3 M 55.6 / 55.0 / 62.1,
// 10 M 223.9 / 219.9 / 242.9. NBX_BVH_NB=1|2|4 forces it (experiments).
static int walk_bodies_per_lane(int prec, uint32_t targets) {
  static const int forced = [] { const char* v = getenv("NBX_BVH_NB"); const int k = v ? atoi(v) : 0; return k == 1 || k == 2 || k == 4 ? k : 0; }();
  if (forced) return forced;
  if (prec == 4) return targets < 1500000u ? 1 : (targets < 2500000u ? 2 : 4);
  return targets < 2500000u ? 1 : 2;
}

// bodies per warp and slot: small problems are bound by the LONGEST warp's chain of dependent steps (about 300 cycles
// each), and the union path of fewer neighbours is shorter (tools/bvh_walk_sim.c, n = 100 k: max 7349 / 5908 / 4827 steps for
// 32 / 16 / 8 bodies per warp) — while there are warp slots to spare, partially filled warps finish sooner. Measured (B200,
// step ms for 32 / 16 / 8 bodies per warp, sibling touch on): float n = 10 k 0.442 / 0.435 / 0.360, 30 k 0.806 / 0.714 /
// 0.688, 100 k 1.536 / 1.598 / 2.008; double 10 k 0.581 / 0.561 / 0.460, 30 k 1.081 / 0.967 / 0.952, 100 k 2.089 / 2.285 /
// 2.991 (tools/exp_bvh_touch.py). A body's arithmetic does not depend on it: results are bit-identical (checked there).
// NBX_BVH_LANES=8|16|32 forces it (experiments).
static uint32_t walk_lanes(int nb, uint32_t targets) {
  const char* v = getenv("NBX_BVH_LANES");
  const int k   = v ? atoi(v) : 0;
  if (k == 8 || k == 16 || k == 32) return uint32_t(k);
  if (nb != 1) return 32u;
  return targets <= 40000u ? 8u : (targets <= 70000u ? 16u : 32u);
}
// sibling sector touch (see touch_sector): on for the one-body-per-lane walk below 250 k targets; NBX_BVH_TOUCH=0|1 forces it
static bool walk_touch(int nb, uint32_t targets) {
  if (nb != 1) return false;
  if (const char* v = getenv("NBX_BVH_TOUCH")) return atoi(v) != 0;
  return targets <= 250000u;
}

template <typename T, int D, bool COUNT>
static int launch_walk(nbx_engine* e, BvhState<T>* s, unsigned long long* stats) {
  const uint32_t nt = e->te - e->tb;
  const T theta     = T(e->cfg.theta);
  const int nb      = walk_bodies_per_lane(e->prec, nt);
  const uint32_t lanes = walk_lanes(nb, nt);
  const auto* xm    = static_cast<const vec4_t<T>*>(e->xm[e->cur]);
  auto* a           = static_cast<vec4_t<T>*>(e->a);
  const unsigned grid = (nt + 4u * lanes * nb - 1) / (4u * lanes * nb);
  PeerOut peers{};
  if (!COUNT && e->peers_ready && e->cfg.world_size > 1) {
    const int k = e->a == e->own_a[0] ? 0 : 1;  // which of its two buffers `a` is right now (the same on every rank)
    for (int r = 0; r < e->cfg.world_size; ++r)
      if (r != e->cfg.rank) peers.p[peers.count++] = e->peer_a[k][r];
  }
#define NBX_WALK(NB_) bvh_force_key_kernel<T, D, NB_, COUNT><<<grid, 128, 0, e->stream>>>(xm, s->rec - 1, e->n, e->tb, e->te, s->levels, theta * theta, T(e->cfg.G), a, stats, peers, lanes)
  if (nb == 1 && walk_touch(nb, nt))
    bvh_force_key_kernel<T, D, 1, COUNT, true><<<grid, 128, 0, e->stream>>>(xm, s->rec - 1, e->n, e->tb, e->te, s->levels, theta * theta, T(e->cfg.G), a, stats, peers, lanes);
  else if (nb == 1) NBX_WALK(1);
  else if (nb == 2) NBX_WALK(2);
  else NBX_WALK(4);
#undef NBX_WALK
  return NBX_OK;
}

template <typename T, int D>
static int force_impl(nbx_engine* e) {
  auto* s = st<T>(e);
  if (!s->built) return fail(NBX_ERR_STATE, "bvh compute_force before build_tree");
  const uint32_t nt = e->te - e->tb;
  if (nt == 0) return NBX_OK;
  const T theta = T(e->cfg.theta);
  static const bool per_thread = [] { const char* v = getenv("NBX_BVH_PER_THREAD"); return v && atoi(v); }();
  if (s->levels <= 27 && !per_thread) NBX_TRY((launch_walk<T, D, false>(e, s, nullptr)));
  else
    bvh_force_kernel<T, D><<<(nt + 127) / 128, 128, 0, e->stream>>>(static_cast<const vec4_t<T>*>(e->xm[e->cur]), s->node_m, s->bw, e->n,
                                                                   e->tb, e->te, s->levels, theta * theta, T(e->cfg.G),
                                                                   static_cast<vec4_t<T>*>(e->a));
  e->launches++;
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

template <typename T, int D>
static int get_bbox_impl(nbx_engine* e, void* xmin, void* xmax) {
  auto* s = st<T>(e);
  Box<T> h;
  NBX_CUDA(cudaMemcpyAsync(&h, s->box, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
  NBX_CUDA(cudaStreamSynchronize(e->stream));
  e->d2h += sizeof(h);
  for (int k = 0; k < D; ++k) {
    if (xmin) static_cast<T*>(xmin)[k] = h.lo[k];
    if (xmax) static_cast<T*>(xmax)[k] = h.hi[k];
  }
  return NBX_OK;
}

template <typename T, int D>
static int get_keys_impl(nbx_engine* e, uint64_t* keys, uint32_t* perm) {
  auto* s = st<T>(e);
  if (!s->sorted) return fail(NBX_ERR_STATE, "no hilbert_sort has run yet");
  if (keys) NBX_CUDA(cudaMemcpyAsync(keys, s->keys, sizeof(uint64_t) * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
  if (perm) NBX_CUDA(cudaMemcpyAsync(perm, s->perm, sizeof(uint32_t) * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
  NBX_CUDA(cudaStreamSynchronize(e->stream));
  e->d2h += (keys ? 8 : 0) * size_t(e->n) + (perm ? 4 : 0) * size_t(e->n);
  return NBX_OK;
}

template <typename T, int D>
static int get_nodes_impl(nbx_engine* e, uint64_t* nnodes, void* node_m, void* bw, void* b) {
  auto* s = st<T>(e);
  if (nnodes) *nnodes = s->nnodes;
  if (!node_m && !bw && !b) return NBX_OK;
  if (!node_m || !bw || !b) return fail(NBX_ERR_INVALID, "bvh_get_nodes: pass all three arrays or none");
  if (!s->built) return fail(NBX_ERR_STATE, "no build_tree has run yet");
  const uint64_t CH = 1u << 20;  // export in chunks through a bounded staging buffer
  T *dm = nullptr, *dw = nullptr, *db = nullptr;
  NBX_CUDA(cudaMalloc(&dm, sizeof(T) * CH * (D + 1)));
  NBX_CUDA(cudaMalloc(&dw, sizeof(T) * CH));
  NBX_CUDA(cudaMalloc(&db, sizeof(T) * CH * 2 * D));
  int rc = NBX_OK;
  for (uint64_t first = 0; first < s->nnodes && rc == NBX_OK; first += CH) {
    uint64_t cnt = std::min<uint64_t>(CH, s->nnodes - first);
    export_nodes_kernel<T, D><<<unsigned((cnt + 255) / 256), 256, 0, e->stream>>>(s->node_m, s->bw, s->lo, s->hi, first, cnt, dm, dw, db);
    e->launches++;
    cudaError_t err = cudaMemcpyAsync(static_cast<T*>(node_m) + first * (D + 1), dm, sizeof(T) * cnt * (D + 1), cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(static_cast<T*>(bw) + first, dw, sizeof(T) * cnt, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(static_cast<T*>(b) + first * 2 * D, db, sizeof(T) * cnt * 2 * D, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) rc = fail(NBX_ERR_CUDA, cudaGetErrorString(err));
    e->d2h += sizeof(T) * cnt * (3 * D + 2);
  }
  cudaFree(dm); cudaFree(dw); cudaFree(db);
  return rc;
}

template <typename T, int D>
static int stats_impl(nbx_engine* e, unsigned long long* dev_stats) {
  auto* s = st<T>(e);
  if (!s->built) return fail(NBX_ERR_STATE, "no build_tree has run yet");
  if (s->levels > 27) return fail(NBX_ERR_INVALID, "traversal stats need n <= 2^27");
  if (e->te > e->tb) NBX_TRY((launch_walk<T, D, true>(e, s, dev_stats)));
  e->launches++;
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

#define BVH_DISPATCH(e, fn, ...)                                                          \
  ((e)->prec == 4 ? ((e)->dim == 2 ? fn<float, 2>(__VA_ARGS__) : fn<float, 3>(__VA_ARGS__)) \
                  : ((e)->dim == 2 ? fn<double, 2>(__VA_ARGS__) : fn<double, 3>(__VA_ARGS__)))

int bvh_create(nbx_engine* e) { return BVH_DISPATCH(e, create_impl, e); }
void bvh_destroy(nbx_engine* e) {
  if (!e->bvh) return;
  BVH_DISPATCH(e, destroy_impl, e);
}
int bvh_bounding_box(nbx_engine* e) {
  PhaseTimer pt(e, PH_BBOX);
  return BVH_DISPATCH(e, bbox_impl, e);
}
int bvh_hilbert_sort(nbx_engine* e) {
  PhaseTimer pt(e, PH_SORT);
  return BVH_DISPATCH(e, sort_impl, e);
}
int bvh_build_tree(nbx_engine* e) {
  PhaseTimer pt(e, PH_MONO);
  return BVH_DISPATCH(e, build_impl, e);
}
int bvh_compute_force(nbx_engine* e) {
  // Peer stores land in the OTHER ranks' acceleration arrays, which their own gather (hilbert_sort) has just written:
  // no rank may start storing before every rank is past its build. Without this barrier a fast rank's results were
  // overwritten by a slower rank's gather — seen on 8 GPUs at n = 50 021 (1 ms steps), not at n = 10 M.
  if (e->cfg.world_size > 1 && e->peers_ready) NBX_TRY(comm_barrier(e));
  {
    PhaseTimer pt(e, PH_TRAVERSE);
    NBX_TRY(BVH_DISPATCH(e, force_impl, e));
  }
  // multi-GPU: a[tb, te) of every rank -> the full acceleration array everywhere (replicated state): either the walk stored
  // its results into the peers' arrays itself and only a barrier is left, or one NCCL all-gather
  if (e->cfg.world_size <= 1) return NBX_OK;
  return e->peers_ready ? comm_barrier(e) : comm_allgather(e, e->a);
}
int bvh_stats(nbx_engine* e, unsigned long long* dev_stats) { return BVH_DISPATCH(e, stats_impl, e, dev_stats); }
// what sort_impl / build_impl do on the host besides enqueueing kernels; a replayed graph step needs the same
void bvh_after_graph_replay(nbx_engine* e) {
  e->cur ^= 1;
  std::swap(e->v, e->v_alt);
  std::swap(e->a, e->a_alt);
  std::swap(e->ao, e->ao_alt);
}
int bvh_walk_width(const nbx_engine* e) {
  const int nb = walk_bodies_per_lane(e->prec, e->te - e->tb);
  return int(walk_lanes(nb, e->te - e->tb)) * nb;
}
int bvh_get_bbox(nbx_engine* e, void* xmin, void* xmax) { return BVH_DISPATCH(e, get_bbox_impl, e, xmin, xmax); }
int bvh_get_keys(nbx_engine* e, uint64_t* keys, uint32_t* perm) { return BVH_DISPATCH(e, get_keys_impl, e, keys, perm); }
int bvh_get_nodes(nbx_engine* e, uint64_t* nnodes, void* node_m, void* bw, void* b) {
  return BVH_DISPATCH(e, get_nodes_impl, e, nnodes, node_m, bw, b);
}

}  // namespace nbx
