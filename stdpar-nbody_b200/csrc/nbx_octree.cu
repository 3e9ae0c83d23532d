// nbx_octree.cu — Barnes-Hut octree (quadtree in 2-D) for sm_100a, built by sorting instead of by locks.
//
// The reference builds the tree with a CAS-locked concurrent insert and a bump allocator (src/octree.h:115-181), then
// climbs from every leaf with an acq_rel latch (src/octree.h:184-224) and walks it stacklessly through parent links
// (src/octree.h:63-71,227-255). Node numbering there depends on the thread schedule; the SET of cells, the child order
// and every monopole do not. This file produces that same set of cells deterministically:
//
//   bounds kernels     : octree::compute_bounds (src/octree.h:93-112)                                   [K6]
//   path_keys_kernel   : each body replays the descent arithmetic of octree::insert (src/octree.h:126-138) on its own —
//                        child bit d = pos[d] > divide[d], divide[d] += (+-1)*side/4, side /= 2 — in the reference's
//                        exact floating-point order, giving a D-bit digit per level, root first                [K7]
//   sort_pairs         : radix sort of the path keys (nbx_sort.cu); sorted order == DFS order of the leaves
//   cells kernels      : a cell of depth d exists iff two adjacent sorted bodies share d digits; cells are emitted in
//                        DFS pre-order into ONE linear array together with the body leaves ("records")     [K5/K7]
//   monopole kernels   : children summed in child order 0..2^D-1, deepest level first (src/octree.h:205-216)   [K8]
//   force_kernel       : walk over the pre-order array: accept -> jump to `next` (end of subtree), open -> p+1 [K9]
//
// Empty children are not stored: in the reference they contribute m=0 at x=0, i.e. exactly +0 to every sum.
// Record = monopole (x,y,z,m) + {next, depth|leaf<<8}: 24 B (float) / 40 B (double), read strictly front to back.
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <type_traits>

#include "nbx_hilbert.cuh"
#include "nbx_internal.cuh"
#include "nbx_math.cuh"

namespace nbx {

namespace {


template <int D>
struct KeyTraits {
  static constexpr int MAXL = D == 3 ? 21 : 32;  // levels that fit a 64-bit key
  static constexpr int BITS = D * MAXL;          // 63 / 64
};

constexpr uint32_t LEAF_FLAG = 0x100u;

template <typename T>
struct Root {
  T side;
  T x;          // root_x = splat(divide)
  uint32_t cells;     // internal cells of the last build
  uint32_t overflow;  // bit 0: bodies not separated within the available levels; bit 1: more cells than capacity
};

template <typename T>
struct OctreeState {
  Root<T>* root      = nullptr;
  T* partial         = nullptr;  // [blocks][2]
  uint32_t nblocks   = 0;
  uint64_t* keys     = nullptr;  // path keys (levels 1..MAXL), entry order
  uint64_t* skeys    = nullptr;  // sorted
  uint64_t* keys_lo  = nullptr;  // deep mode: levels MAXL+1..2*MAXL, entry order / scratch
  uint64_t* skeys_lo = nullptr;  // deep mode: sorted low words (all zero in shallow mode)
  uint64_t* keys_tmp = nullptr;  // deep mode scratch
  uint32_t* perm_tmp = nullptr;  // deep mode scratch
  bool deep          = false;    // the last build needed more than MAXL levels
  uint32_t* perm     = nullptr;  // sorted slot -> body id
  uint32_t* delta    = nullptr;  // [n] common-prefix levels of sorted neighbours s, s+1 (MAXL+1.. see kernel)
  uint32_t* cnt      = nullptr;  // [n+1] cells starting at sorted body s -> exclusive scan in place (cell_base)
  uint32_t* blocksum = nullptr;
  uint32_t ccap      = 0;        // internal-cell capacity = max(n, 256)  (reference: max(2^D n, 1000) nodes, system.h:29)
  uint32_t cap       = 0;        // record capacity = n + ccap + 1
  vec4_t<T>* mono    = nullptr;  // [cap]
  uint2* meta        = nullptr;  // [cap] {next, depth | leaf<<8}
  uint32_t* rec_body = nullptr;  // [cap] first sorted body of the node (for the canonical path code)
  uint32_t* cell_pos = nullptr;  // [n] record index of internal cell c
  uint32_t* cells_by_depth = nullptr;  // [n] record indices of the cells grouped by depth (for the level-wise up-pass)
  uint32_t* depth_count = nullptr;     // [130] cells per depth -> exclusive offsets [0..128], cursor copy at +...
  uint32_t* depth_cursor = nullptr;    // [129]
  vec4_t<T>* a_sorted = nullptr; // [n_pad] accelerations in sorted-slot order
  uint64_t* hkeys     = nullptr; // [n] coarse Hilbert index of sorted slot s (walk_order_keys_kernel)
  uint32_t* order     = nullptr; // [n] lane slot t -> sorted slot: the targets in Hilbert order (nullptr: path order)
  bool hilbert_targets = true;
  T* thr_table        = nullptr; // [2][132] acceptance thresholds per depth on d2 / on dx (threshold_table_kernel)
  bool built = false;
};

// ---- K6 bounds (octree.h:93-112): scalar min/max over every coordinate, identity 0 --------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(256) bounds_partial_kernel(const vec4_t<T>* __restrict__ xm, uint32_t n, T* partial) {
  T lo = 0, hi = 0;
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    vec4_t<T> b = xm[i];
    T mn = fmin(b.x, b.y), mx = fmax(b.x, b.y);
    if (D == 3) { mn = fmin(mn, b.z); mx = fmax(mx, b.z); }
    lo = fmin(lo, mn);
    hi = fmax(hi, mx);
  }
  __shared__ T red[8][2];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = lo; red[threadIdx.x >> 5][1] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { lo = fmin(lo, red[w][0]); hi = fmax(hi, red[w][1]); }
    partial[blockIdx.x * 2]     = lo;
    partial[blockIdx.x * 2 + 1] = hi;
  }
}

template <typename T>
__global__ void bounds_final_kernel(const T* partial, uint32_t nblocks, Root<T>* root) {
  const int lane = threadIdx.x;  // one warp
  T lo = 0, hi = 0;
  for (uint32_t b = lane; b < nblocks; b += 32) { lo = fmin(lo, partial[2 * b]); hi = fmax(hi, partial[2 * b + 1]); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if (lane != 0) return;
  hi = add_rn(hi, T(1));                          // max_size += 1
  lo = sub_rn(lo, T(1));                          // min_size -= 1
  root->x        = div_rn(add_rn(hi, lo), T(2));  // divide
  root->side     = sub_rn(hi, lo);
  root->overflow = 0;
  root->cells    = 0;
}

// ---- K7 path keys -------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(256) path_keys_kernel(const vec4_t<T>* __restrict__ xm, uint32_t n,
                                                        const Root<T>* __restrict__ root, uint64_t* __restrict__ keys,
                                                        uint64_t* __restrict__ keys_lo /* NULL: first MAXL levels only */) {
  uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  vec4_t<T> b = xm[i];
  const T pos[3] = {b.x, b.y, b.z};
  T divide[3]    = {root->x, root->x, root->x};
  T side         = root->side;
  uint64_t key   = 0;
#pragma unroll 1
  for (int level = 0; level < KeyTraits<D>::MAXL; ++level) {
    const T half = div_rn(side, T(4));  // "/4: /2 for the new side, /2 for half of it" (octree.h:126-127)
    uint32_t cp  = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const bool gt = pos[k] > divide[k];
      cp |= uint32_t(gt) << k;
      divide[k] = add_rn(divide[k], gt ? half : -half);  // (2*gt - 1) * half
    }
    side = div_rn(side, T(2));
    key  = (key << D) | cp;
  }
  keys[i] = key;
  if (keys_lo == nullptr) return;
  key = 0;  // the same descent, MAXL more levels
#pragma unroll 1
  for (int level = 0; level < KeyTraits<D>::MAXL; ++level) {
    const T half = div_rn(side, T(4));
    uint32_t cp  = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const bool gt = pos[k] > divide[k];
      cp |= uint32_t(gt) << k;
      divide[k] = add_rn(divide[k], gt ? half : -half);
    }
    side = div_rn(side, T(2));
    key  = (key << D) | cp;
  }
  keys_lo[i] = key;
}

__global__ void __launch_bounds__(256) gather_u64_kernel(const uint64_t* __restrict__ src, const uint32_t* __restrict__ idx,
                                                         uint32_t n, uint64_t* __restrict__ dst) {
  uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// ---- cells --------------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ uint32_t common_levels(uint64_t a, uint64_t b) {
  uint64_t x = a ^ b;
  if (x == 0) return KeyTraits<D>::MAXL;
  int lead = __clzll((long long)x) - (64 - KeyTraits<D>::BITS);
  return uint32_t(lead / D);
}

// two-word keys: levels 1..MAXL in `hi`, MAXL+1..2*MAXL in `lo` (lo == NULL: shallow mode)
template <int D>
__device__ __forceinline__ uint32_t common_levels2(const uint64_t* hi, const uint64_t* lo, uint32_t a, uint32_t b) {
  uint32_t c = common_levels<D>(hi[a], hi[b]);
  if (c == KeyTraits<D>::MAXL && lo) c += common_levels<D>(lo[a], lo[b]);
  return c;
}
// the first `depth` digits of sorted body i, as (word, value): equal prefixes <=> equal pairs
template <int D>
__device__ __forceinline__ bool same_prefix(const uint64_t* hi, const uint64_t* lo, uint32_t a, uint32_t b, uint32_t depth) {
  constexpr int MAXL = KeyTraits<D>::MAXL;
  if (depth == 0) return true;
  if (depth <= MAXL) {
    const int shift = D * (MAXL - int(depth));
    return shift >= 64 ? true : (hi[a] >> shift) == (hi[b] >> shift);
  }
  if (hi[a] != hi[b]) return false;
  const int shift = D * (2 * MAXL - int(depth));
  return shift >= 64 ? true : (lo[a] >> shift) == (lo[b] >> shift);
}

// delta[s] = levels shared by sorted bodies s and s+1 (s < n-1); cnt[s] = number of cells whose first body is s
template <int D>
__global__ void __launch_bounds__(256) delta_kernel(const uint64_t* __restrict__ skeys, const uint64_t* __restrict__ skeys_lo,
                                                    uint32_t n, uint32_t* delta, uint32_t* cnt, uint32_t* overflow) {
  uint32_t s = blockIdx.x * 256 + threadIdx.x;
  if (s > n) return;
  if (s == n) { cnt[n] = 0; return; }
  // shared levels with the next / previous body; "-1" is encoded by computing with +1 offsets
  uint32_t dn = s + 1 < n ? common_levels2<D>(skeys, skeys_lo, s, s + 1) + 1 : 0;  // delta_s + 1   (0 = none)
  uint32_t dp = s > 0 ? common_levels2<D>(skeys, skeys_lo, s - 1, s) + 1 : 0;      // delta_{s-1} + 1
  const uint32_t maxl = skeys_lo ? 2 * KeyTraits<D>::MAXL : KeyTraits<D>::MAXL;
  if (dn == maxl + 1) atomicOr(overflow, 1u);  // bit 0: not separated within the available levels
  delta[s] = dn;
  cnt[s]   = dn > dp ? dn - dp : 0;
}

// generic exclusive scan of u32 (3 kernels), in place; data has `count` entries
constexpr int SC_TILE = 4096;
__global__ void __launch_bounds__(1024) scan_reduce_kernel(const uint32_t* data, uint32_t count, uint32_t* blocksum) {
  __shared__ uint32_t ws[32];
  uint32_t base = blockIdx.x * SC_TILE, v = 0;
  for (int k = 0; k < 4; ++k) {
    uint32_t i = base + k * 1024 + threadIdx.x;
    if (i < count) v += data[i];
  }
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = ws[threadIdx.x];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0) blocksum[blockIdx.x] = v;
  }
}
__global__ void __launch_bounds__(1024) scan_blocksums_kernel(uint32_t* blocksum, uint32_t nb, uint32_t* total_out) {
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < nb; base += 1024) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < nb ? blocksum[i] : 0, s = v;
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, s, off);
      if (lane >= off) s += t;
    }
    if (lane == 31) ws[warp] = s;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = ws[lane];
      for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, w, off);
        if (lane >= off) w += t;
      }
      ws[lane] = w;
    }
    __syncthreads();
    uint32_t ex = carry + (warp ? ws[warp - 1] : 0) + (s - v);
    if (i < nb) blocksum[i] = ex;
    __syncthreads();
    if (threadIdx.x == 1023) carry = ex + v;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}
__global__ void __launch_bounds__(1024) scan_apply_kernel(uint32_t* data, uint32_t count, const uint32_t* blocksum) {
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = blocksum[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = blockIdx.x * SC_TILE;
  for (int k = 0; k < 4; ++k) {
    uint32_t i = base + k * 1024 + threadIdx.x;
    uint32_t v = i < count ? data[i] : 0, s = v;
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, s, off);
      if (lane >= off) s += t;
    }
    if (lane == 31) ws[warp] = s;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = ws[lane];
      for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, w, off);
        if (lane >= off) w += t;
      }
      ws[lane] = w;
    }
    __syncthreads();
    uint32_t ex = carry + (warp ? ws[warp - 1] : 0) + (s - v);
    if (i < count) data[i] = ex;
    __syncthreads();
    if (threadIdx.x == 1023) carry = ex + v;
    __syncthreads();
  }
}

template <typename T, int D>
__device__ __forceinline__ void emit_records_body(const uint64_t* __restrict__ skeys, const uint64_t* __restrict__ skeys_lo,
                                                  const uint32_t* __restrict__ perm, const vec4_t<T>* __restrict__ xm, uint32_t n,
                                                  const uint32_t* __restrict__ delta, const uint32_t* __restrict__ cell_base,
                                                  uint32_t cap, uint32_t ccap, vec4_t<T>* mono, uint2* meta, uint32_t* rec_body,
                                                  uint32_t* cell_pos, Root<T>* root, uint32_t* dhist);

// One thread per sorted body s: emits its leaf record and the records of every cell that starts at s.
// Record index of cell (s, depth d) = s + cell_id, cell_id = cell_base[s] + (d - first depth); of leaf s = s + cell_base[s+1].
// `next` of a cell = record index right after its last body e: (e + 1) + cell_base[e + 1].
template <typename T, int D>
__global__ void __launch_bounds__(256) emit_records_kernel(const uint64_t* __restrict__ skeys, const uint64_t* __restrict__ skeys_lo,
                                                           const uint32_t* __restrict__ perm,
                                                           const vec4_t<T>* __restrict__ xm, uint32_t n,
                                                           const uint32_t* __restrict__ delta, const uint32_t* __restrict__ cell_base,
                                                           uint32_t cap, uint32_t ccap, vec4_t<T>* mono, uint2* meta, uint32_t* rec_body,
                                                           uint32_t* cell_pos, Root<T>* root, uint32_t* depth_count) {
  __shared__ uint32_t dhist[130];
  for (int q = threadIdx.x; q < 130; q += 256) dhist[q] = 0;
  __syncthreads();
  emit_records_body<T, D>(skeys, skeys_lo, perm, xm, n, delta, cell_base, cap, ccap, mono, meta, rec_body, cell_pos, root, dhist);
  __syncthreads();
  for (int q = threadIdx.x; q < 130; q += 256)
    if (dhist[q]) atomicAdd(&depth_count[q], dhist[q]);
}

template <typename T, int D>
__device__ __forceinline__ void emit_records_body(const uint64_t* __restrict__ skeys, const uint64_t* __restrict__ skeys_lo,
                                                  const uint32_t* __restrict__ perm, const vec4_t<T>* __restrict__ xm, uint32_t n,
                                                  const uint32_t* __restrict__ delta, const uint32_t* __restrict__ cell_base,
                                                  uint32_t cap, uint32_t ccap, vec4_t<T>* mono, uint2* meta, uint32_t* rec_body,
                                                  uint32_t* cell_pos, Root<T>* root, uint32_t* dhist) {
  uint32_t s = blockIdx.x * 256 + threadIdx.x;
  if (s >= n) return;
  const uint32_t dn = delta[s];                      // delta_s + 1
  const uint32_t dp = s > 0 ? delta[s - 1] : 0;      // delta_{s-1} + 1
  const uint32_t cb = cell_base[s], cb_next = cell_base[s + 1];
  if (s == n - 1) {
    root->cells = cb_next;
    if (cb_next > ccap) atomicOr(&root->overflow, 2u);  // bit 1: more internal cells than capacity
  }
  // leaf: depth = deepest shared level + 1
  const uint32_t leaf_depth = (dn > dp ? dn : dp);  // max(delta_s, delta_{s-1}) + 1, and 0 for a single body
  const uint32_t lpos       = s + cb_next;
  if (lpos < cap) {
    mono[lpos]     = xm[perm[s]];
    meta[lpos]     = make_uint2(lpos + 1, leaf_depth | LEAF_FLAG);
    rec_body[lpos] = s;
  }
  if (dn <= dp) return;
  for (uint32_t q = 0; q < dn - dp; ++q) {
    const uint32_t depth = dp + q;  // cells at depths delta_{s-1}+1 .. delta_s  (dp = delta_{s-1}+1)
    // last body e sharing `depth` digits with s: binary search on the sorted keys
    uint32_t lo = s, hi = n - 1;  // invariant: lo shares the prefix
    while (lo < hi) {
      uint32_t mid = lo + (hi - lo + 1) / 2;
      if (same_prefix<D>(skeys, skeys_lo, s, mid, depth)) lo = mid;
      else hi = mid - 1;
    }
    const uint32_t e   = lo;
    const uint32_t cid = cb + q;
    const uint32_t pos = s + cid;
    if (pos < cap && cid < ccap) {
      meta[pos]     = make_uint2((e + 1) + cell_base[e + 1], depth);
      rec_body[pos] = s;
      cell_pos[cid] = pos;
      atomicAdd(&dhist[depth], 1u);
    }
  }
}

// ---- K8 monopoles: one launch per depth, deepest first (octree.h:205-216: m = sum m_c ; x = sum m_c*x_c / m) -------------
// cells are first grouped by depth (counting sort: depth_count -> offsets -> cells_by_depth) so that every level launch only
// touches its own cells.
__global__ void depth_offsets_kernel(uint32_t* depth_count, uint32_t* depth_cursor) {
  if (threadIdx.x != 0) return;
  uint32_t run = 0;
  for (int d = 0; d < 129; ++d) {
    uint32_t c      = d < 128 ? depth_count[d] : 0;
    depth_count[d]  = run;  // exclusive offsets
    depth_cursor[d] = run;
    run += c;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) group_cells_kernel(const uint32_t* __restrict__ cell_pos, const Root<T>* __restrict__ root,
                                                          const uint2* __restrict__ meta, uint32_t cap, uint32_t* depth_cursor,
                                                          uint32_t* __restrict__ cells_by_depth) {
  __shared__ uint32_t cnt[130], base[130];
  for (int q = threadIdx.x; q < 130; q += 256) cnt[q] = 0;
  __syncthreads();
  const uint32_t c = blockIdx.x * 256 + threadIdx.x;
  uint32_t p = 0xffffffffu, d = 0, local = 0;
  if (c < root->cells) {
    p = cell_pos[c];
    if (p < cap) {
      d     = meta[p].y & 0xff;
      local = atomicAdd(&cnt[d], 1u);  // order inside a depth does not matter: every cell is computed independently
    } else {
      p = 0xffffffffu;
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < 130; q += 256)
    if (cnt[q]) base[q] = atomicAdd(&depth_cursor[q], cnt[q]);
  __syncthreads();
  if (p != 0xffffffffu) cells_by_depth[base[d] + local] = p;
}
template <typename T, int D>
__device__ __forceinline__ void monopole_cell(uint32_t p, uint32_t cap, vec4_t<T>* mono, const uint2* __restrict__ meta) {
  const uint2 me = meta[p];
  T m = 0, x = 0, y = 0, z = 0;
  uint32_t q = p + 1;
  while (q < me.x && q < cap) {  // children in child order
    const vec4_t<T> cm = mono[q];
    m = add_rn(m, cm.w);
    x = add_rn(x, mul_rn(cm.w, cm.x));
    y = add_rn(y, mul_rn(cm.w, cm.y));
    if (D == 3) z = add_rn(z, mul_rn(cm.w, cm.z));
    q = meta[q].x;
  }
  mono[p] = make_v4<T>(div_rn(x, m), div_rn(y, m), D == 3 ? div_rn(z, m) : T(0), m);
}
template <typename T, int D>
__global__ void __launch_bounds__(256) monopole_level_kernel(const uint32_t* __restrict__ cells_by_depth, const uint32_t* __restrict__ depth_off,
                                                             uint32_t depth, uint32_t cap, vec4_t<T>* mono, const uint2* __restrict__ meta) {
  const uint32_t first = depth_off[depth], count = depth_off[depth + 1] - first;
  uint32_t c = blockIdx.x * 256 + threadIdx.x;
  if (c >= count) return;
  monopole_cell<T, D>(cells_by_depth[first + c], cap, mono, meta);
}
// depths `top` .. 0 in ONE launch when each holds at most 1024 cells (the top of every tree): a single CTA walks up
template <typename T, int D>
__global__ void __launch_bounds__(1024) monopole_top_kernel(const uint32_t* __restrict__ cells_by_depth, const uint32_t* __restrict__ depth_off,
                                                            int top, uint32_t cap, vec4_t<T>* mono, const uint2* __restrict__ meta) {
  for (int depth = top; depth >= 0; --depth) {
    const uint32_t first = depth_off[depth], count = depth_off[depth + 1] - first;
    if (threadIdx.x < count) monopole_cell<T, D>(cells_by_depth[first + threadIdx.x], cap, mono, meta);
    __syncthreads();
  }
}

// ---- target order of the walk ------------------------------------------------------------------------------------------
// The sorted slots are in the tree's DFS order = Z (Morton) order of the path keys, and a Z curve jumps: 32 consecutive
// slots regularly straddle the boundary of a big cell and then hold two far-apart clumps, whose union walk is nearly the
// sum of both. The lanes of a warp are therefore assigned along a HILBERT curve through the same cells (consecutive
// cells always share a face): the top HB levels of the sorted path key are de-interleaved into cell coordinates, run
// through Skilling's transform ("Programming the Hilbert curve", AIP Conf. Proc. 707, 2004) and re-interleaved; a stable
// sort of that coarse index (HB*D <= 32 bits: 4 radix passes) gives order[t] = sorted slot of lane slot t. Only the
// assignment of targets to lanes changes — every body still performs its own sequence of tests on the same records.
// HB = 10 levels in 3-D (measured at n = 10 M: 8 levels gain nothing, 13 and 16 gain no more than 10 and sort longer).
template <int D>
__global__ void __launch_bounds__(256) walk_order_keys_kernel(const uint64_t* __restrict__ skeys, uint32_t n, int HB,
                                                              uint64_t* __restrict__ hkeys) {
  const uint32_t s = blockIdx.x * 256 + threadIdx.x;
  if (s >= n) return;
  const uint64_t key = skeys[s];
  uint32_t X[D];
#pragma unroll
  for (int k = 0; k < D; ++k) X[k] = 0;
  for (int l = 0; l < HB; ++l) {  // digit of level l (root first): bit k = axis k (path_keys_kernel)
    const uint32_t digit = uint32_t(key >> (D * (KeyTraits<D>::MAXL - 1 - l))) & ((1u << D) - 1);
#pragma unroll
    for (int k = 0; k < D; ++k) X[k] |= ((digit >> k) & 1u) << (HB - 1 - l);
  }
  hkeys[s] = hilbert_index<D>(X, HB);
}

// ---- K9 traversal ---------------------------------------------------------------------------------------------------
// octree.h:227-255: dx = sqrt(dist2)+eps ; accept when leaf or side/dx < theta ; a += m*(xj-x)/dx^3.
// The reference's decision is a MONOTONE function of dist2 (sqrt.rn and +eps are non-decreasing, side/dx is
// non-increasing in dx), so for every depth there is exactly one value tau(depth) with
//     side(depth) / (sqrt(d2) + eps) < theta   <=>   d2 >= tau(depth)        for every representable d2 >= 0.
// threshold_table_kernel finds tau by bisection over the bit patterns of T with the reference's own IEEE operations
// (sqrt.rn, add.rn, div.rn), once per step; the walk then compares the reference-order dist2 with the table.
// exact halvings of the root side: side(depth) = root_side * 2^-depth
__device__ __forceinline__ float side_at(float root_side, uint32_t depth) { return root_side * __int_as_float(int(127 - depth) << 23); }
__device__ __forceinline__ double side_at(double root_side, uint32_t depth) {
  return root_side * __longlong_as_double((long long)(1023 - depth) << 52);
}

// WARP-COOPERATIVE walk: one lane per SORTED slot t (body perm[t]); the 32 lanes of a warp are neighbours along the
// tree's DFS order. The warp walks the UNION of its lanes' paths with ONE position `p` (warp-uniform => every record
// load is a single broadcast sector instead of up to 32 divergent ones), while each lane keeps the reference's per-body
// semantics exactly: a lane that accepts node p while another lane needs it opened simply sleeps until the walk leaves
// that subtree (`resume` = the record index where it wakes up). Interaction sets are identical to the per-body walk.
// `s_table` (threshold_table_kernel) holds tau(depth) and 0 in the leaf slot [128] (always accepted), so the whole
// acceptance test is one shared-memory load and one compare on the reference-order dist2 — the SAME decision as the
// reference's `side/dx < theta` for every input, not just away from the threshold; the square root is only needed for the
// accumulation.
template <typename T, int D, bool COUNT = false>
__global__ void __launch_bounds__(128) octree_force_kernel(const vec4_t<T>* __restrict__ mono, const uint2* __restrict__ meta,
                                                           const Root<T>* __restrict__ root, const uint32_t* __restrict__ cell_base,
                                                           const uint32_t* __restrict__ order, uint32_t n, uint32_t tb, uint32_t te,
                                                           const T* __restrict__ s_table, T c, vec4_t<T>* __restrict__ a_sorted,
                                                           unsigned long long* stats, const uint32_t lanes) {
  __shared__ T tab[132];
  for (uint32_t d = threadIdx.x; d < 132; d += blockDim.x) tab[d] = s_table[d];
  __syncthreads();
  unsigned long long n_visit = 0, n_take = 0, n_step = 0;  // COUNT only
  // lanes: bodies per warp (32, or 16 / 8 for small problems: only the first `lanes` lanes carry bodies, see oct_walk_lanes)
  const uint32_t t    = tb + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * lanes + (threadIdx.x & 31u);
  const bool valid    = (threadIdx.x & 31u) < lanes && t < te;
  const uint32_t nrec = n + root->cells;
  const uint32_t tt   = order ? order[valid ? t : tb] : (valid ? t : tb);  // sorted slot of this lane's body
  const vec4_t<T> xs  = mono[tt + cell_base[tt + 1]];  // own leaf record = own position
  T ax = 0, ay = 0, az = 0;
  uint32_t resume = valid ? 0u : 0xffffffffu;  // first record this lane still has to look at
  uint32_t p      = 0;
  while (p < nrec) {
    const vec4_t<T> nm = mono[p];  // warp-uniform address
    const uint2 me     = meta[p];
    const T tau        = tab[min(me.y, 128u)];
    // dist2 in the reference's order, unfused (vec.h:232-240; (xj - x)^2 == (x - xj)^2 exactly)
    const T dx_ = sub_rn(nm.x, xs.x), dy_ = sub_rn(nm.y, xs.y), dz_ = D == 3 ? sub_rn(nm.z, xs.z) : T(0);
    T d2 = add_rn(mul_rn(dx_, dx_), mul_rn(dy_, dy_));
    if (D == 3) d2 = add_rn(d2, mul_rn(dz_, dz_));
    const T dx      = dist_eps(d2);
    const bool act  = p >= resume;
    const bool take = d2 >= tau;
    if (COUNT) { n_visit += act; n_take += act && take; n_step += 1; }
    if (act && take) {
      const T s = nm.w * inv_cube(dx);
      ax = fma(dx_, s, ax);
      ay = fma(dy_, s, ay);
      if (D == 3) az = fma(dz_, s, az);
      resume = me.x;
    }
    p = __any_sync(0xffffffffu, act && !take) ? p + 1 : me.x;
  }
  if (COUNT) {
    atomicAdd(&stats[0], n_visit);
    atomicAdd(&stats[1], n_take);
    if ((threadIdx.x & 31) == 0) atomicAdd(&stats[2], n_step);
    return;
  }
  if (valid) a_sorted[t] = make_v4<T>(c * ax, c * ay, D == 3 ? c * az : T(0), T(0));
}

// ---- K9, double ------------------------------------------------------------------------------------------------------------
// The walk's critical path (load -> d2 -> decision -> vote -> next load) needs no square root at all: the decision is
// one compare against tau(depth), and the root / reciprocal are only evaluated for the accumulation, off the critical
// path. d2 keeps its fused form here (two FMAs, + 1e-300 so that the rsqrt seed never sees 0): against the reference's
// unfused dist2 it can move by an ulp, so a decision could only differ for a d2 within 2^-52 of tau — the device's test
// counts equal the oracle's in every test. (The float walk evaluates the reference-order dist2 and is exact.)
// m / (sqrt(d2) + eps)^3
__device__ __forceinline__ float mass_inv_cube(float m, float d2) {
  if (d2 < 1e-6f) return m * inv_cube(dist_eps(d2));  // (eps*y)^2 no longer negligible: coincident / self
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d2));  // MUFU.RSQ
  const float yc = fmaf(-FLT_EPSILON * y, y, y);            // 1/(s + eps) = y (1 - eps y + O((eps y)^2))
  return (m * yc) * (yc * yc);
}
__device__ __forceinline__ double mass_inv_cube(double m, double d2) {
  if (__double2hiint(d2) < 0x3ddb7cdf) return m * inv_cube(dist_eps_pos(d2));  // d2 < ~1e-10 (d2 > 0: the high word orders it)
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d2));  // MUFU.RSQ64H, 2^-20
  // y0^3 (1 - e)^(-3/2) with e = 1 - d2 y0^2 - 2 eps y0 (the eps term turns 1/s^3 into 1/(s + eps)^3 to first order)
  const double t  = d2 * y0;
  double e        = fma(-t, y0, 1.0);
  e               = fma(-2.0 * DBL_EPSILON, y0, e);
  const double c2 = e * fma(1.875, e, 1.5);
  const double w3 = (m * y0) * (y0 * y0);
  return fma(w3, c2, w3);
}

// Per-depth acceptance table tau(depth) (see "K9 traversal" above): the smallest d2 >= 0 — as a bit pattern of T — for
// which the reference's predicate side/(sqrt(d2)+eps) < theta holds, found by bisection with IEEE operations. Both walks
// use it: thr_table[d] for d < 128, thr_table[128..131] = 0 (leaf slot: always accepted); +inf when no d2 is accepted
// (theta <= 0). The second half [132 + d] keeps side(d)/theta (rounded from double) for diagnostics.
__device__ __forceinline__ bool ref_accepts(float side, float d2, float theta) {
  return __fdiv_rn(side, __fadd_rn(__fsqrt_rn(d2), FLT_EPSILON)) < theta;
}
__device__ __forceinline__ bool ref_accepts(double side, double d2, double theta) {
  return __ddiv_rn(side, __dadd_rn(__dsqrt_rn(d2), DBL_EPSILON)) < theta;
}
__device__ __forceinline__ float from_bits(uint32_t b, float) { return __uint_as_float(b); }
__device__ __forceinline__ double from_bits(uint64_t b, double) { return __longlong_as_double((long long)b); }

template <typename T>
__global__ void threshold_table_kernel(const Root<T>* __restrict__ root, T theta, T* __restrict__ thr_table) {
  using bits_t = typename std::conditional<sizeof(T) == 4, uint32_t, uint64_t>::type;
  const uint32_t d = threadIdx.x;
  if (d >= 132) return;
  T tau = T(0);
  double S = -1.0;
  if (d < 128) {
    const T side = side_at(root->side, d);
    S            = theta > T(0) ? double(side) / double(theta) : double(INFINITY);
    const bits_t inf_bits = sizeof(T) == 4 ? bits_t(0x7f800000u) : bits_t(0x7ff0000000000000ull);
    if (!ref_accepts(side, from_bits(inf_bits, T(0)), theta)) {
      tau = from_bits(inf_bits, T(0));  // nothing is accepted (theta <= 0): finite d2 >= +inf is never true
    } else if (ref_accepts(side, T(0), theta)) {
      tau = T(0);
    } else {
      bits_t lo = 0, hi = inf_bits;  // predicate false at lo, true at hi; non-negative values order like their bit patterns
      while (hi - lo > 1) {
        const bits_t mid = lo + (hi - lo) / 2;
        if (ref_accepts(side, from_bits(mid, T(0)), theta)) hi = mid;
        else lo = mid;
      }
      tau = from_bits(hi, T(0));
    }
  }
  thr_table[d]       = tau;
  thr_table[132 + d] = T(S);
}

template <typename T, int D, bool COUNT = false>
__global__ void __launch_bounds__(128) octree_force_thr_kernel(const vec4_t<T>* __restrict__ mono, const uint2* __restrict__ meta,
                                                               const Root<T>* __restrict__ root, const uint32_t* __restrict__ cell_base,
                                                               const uint32_t* __restrict__ order, uint32_t n, uint32_t tb, uint32_t te,
                                                               const T* __restrict__ thr_table, T c, vec4_t<T>* __restrict__ a_sorted,
                                                               unsigned long long* stats, const uint32_t lanes) {
  __shared__ T tab[132];  // tau(depth) ; [128] = 0, leaf: always accepted
  for (uint32_t d = threadIdx.x; d < 132; d += blockDim.x) tab[d] = thr_table[d];
  __syncthreads();
  unsigned long long n_visit = 0, n_take = 0, n_step = 0;  // COUNT only
  // lanes: bodies per warp (32, or 16 / 8 for small problems: only the first `lanes` lanes carry bodies, see oct_walk_lanes)
  const uint32_t t    = tb + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * lanes + (threadIdx.x & 31u);
  const bool valid    = (threadIdx.x & 31u) < lanes && t < te;
  const uint32_t nrec = n + root->cells;
  const uint32_t tt   = order ? order[valid ? t : tb] : (valid ? t : tb);  // sorted slot of this lane's body
  const vec4_t<T> xs  = mono[tt + cell_base[tt + 1]];  // own leaf record = own position
  T ax = 0, ay = 0, az = 0;
  uint32_t resume = valid ? 0u : 0xffffffffu;  // first record this lane still has to look at
  uint32_t p      = 0;
  while (p < nrec) {
    const vec4_t<T> nm = mono[p];  // warp-uniform address
    const uint2 me     = meta[p];
    const T thr        = tab[min(me.y, 128u)];
    const T dx_ = nm.x - xs.x, dy_ = nm.y - xs.y, dz_ = D == 3 ? nm.z - xs.z : T(0);
    T d2 = fma(dy_, dy_, sq_plus_tiny(dx_));
    if (D == 3) d2 = fma(dz_, dz_, d2);
    const bool take = d2 >= thr;
    const bool act  = p >= resume;
    if (COUNT) { n_visit += act; n_take += act && take; n_step += 1; }
    if (act && take) {
      const T s = mass_inv_cube(nm.w, d2);
      ax = fma(dx_, s, ax);
      ay = fma(dy_, s, ay);
      if (D == 3) az = fma(dz_, s, az);
      resume = me.x;
    }
    p = __any_sync(0xffffffffu, act && !take) ? p + 1 : me.x;
  }
  if (COUNT) {
    atomicAdd(&stats[0], n_visit);
    atomicAdd(&stats[1], n_take);
    if ((threadIdx.x & 31) == 0) atomicAdd(&stats[2], n_step);
    return;
  }
  if (valid) a_sorted[t] = make_v4<T>(c * ax, c * ay, D == 3 ? c * az : T(0), T(0));
}

// a[perm[order[t]]] = a_sorted[t]   (a_sorted is in lane-slot order)
template <typename T>
__global__ void __launch_bounds__(256) unsort_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ order,
                                                     const vec4_t<T>* __restrict__ a_sorted, uint32_t n, vec4_t<T>* __restrict__ a) {
  uint32_t t = blockIdx.x * 256 + threadIdx.x;
  if (t < n) a[perm[order ? order[t] : t]] = a_sorted[t];
}

template <typename T, int D>
__global__ void export_canonical_kernel(const vec4_t<T>* mono, const uint2* meta, const uint32_t* rec_body,
                                        const uint64_t* skeys, uint32_t nrec, uint32_t* depth, uint64_t* path, uint32_t* kind,
                                        T* out_m) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nrec) return;
  const uint2 me   = meta[p];
  const uint32_t d = me.y & 0xff;
  depth[p]         = d;
  kind[p]          = (me.y & LEAF_FLAG) ? 1u : 0u;
  const int shift  = D * (KeyTraits<D>::MAXL - int(d));
  // the u64 path code only holds MAXL levels: deeper nodes (two-word keys) export ~0
  path[p]          = d == 0 ? 0 : (shift < 0 ? ~0ull : (shift >= 64 ? 0 : skeys[rec_body[p]] >> shift));
  vec4_t<T> m      = mono[p];
  T* om            = out_m + size_t(p) * (D + 1);
  om[0] = m.x; om[1] = m.y;
  if (D == 3) om[2] = m.z;
  om[D] = m.w;
}

template <typename T>
OctreeState<T>* st(nbx_engine* e) { return static_cast<OctreeState<T>*>(e->octree); }

}  // namespace

template <typename T, int D>
static int create_impl(nbx_engine* e) {
  auto* s    = new OctreeState<T>();
  e->octree  = s;
  const size_t n = e->n;
  s->nblocks = std::min<uint32_t>((e->n + 255) / 256, uint32_t(e->sm_count) * 8);
  s->ccap    = std::max<uint32_t>(e->n, 256u);
  s->cap     = uint32_t(std::min<uint64_t>(uint64_t(n) + s->ccap + 1, 0xfffffff0ull));
  NBX_CUDA(cudaMalloc(&s->root, sizeof(Root<T>)));
  NBX_CUDA(cudaMemsetAsync(s->root, 0, sizeof(Root<T>), e->stream));
  NBX_CUDA(cudaMalloc(&s->partial, sizeof(T) * 2 * s->nblocks));
  NBX_CUDA(cudaMalloc(&s->keys, sizeof(uint64_t) * n));
  NBX_CUDA(cudaMalloc(&s->skeys, sizeof(uint64_t) * n));
  NBX_CUDA(cudaMalloc(&s->keys_lo, sizeof(uint64_t) * n));
  NBX_CUDA(cudaMalloc(&s->skeys_lo, sizeof(uint64_t) * n));
  NBX_CUDA(cudaMalloc(&s->keys_tmp, sizeof(uint64_t) * n));
  NBX_CUDA(cudaMalloc(&s->perm_tmp, sizeof(uint32_t) * n));
  NBX_CUDA(cudaMalloc(&s->perm, sizeof(uint32_t) * n));
  NBX_CUDA(cudaMalloc(&s->delta, sizeof(uint32_t) * n));
  NBX_CUDA(cudaMalloc(&s->cnt, sizeof(uint32_t) * (n + 1)));
  NBX_CUDA(cudaMalloc(&s->blocksum, sizeof(uint32_t) * ((n + 1 + SC_TILE - 1) / SC_TILE + 1)));
  NBX_CUDA(cudaMalloc(&s->mono, sizeof(vec4_t<T>) * s->cap));
  NBX_CUDA(cudaMalloc(&s->meta, sizeof(uint2) * s->cap));
  NBX_CUDA(cudaMalloc(&s->rec_body, sizeof(uint32_t) * s->cap));
  NBX_CUDA(cudaMalloc(&s->cell_pos, sizeof(uint32_t) * s->ccap));
  NBX_CUDA(cudaMalloc(&s->cells_by_depth, sizeof(uint32_t) * s->ccap));
  NBX_CUDA(cudaMalloc(&s->depth_count, sizeof(uint32_t) * 130));
  NBX_CUDA(cudaMalloc(&s->depth_cursor, sizeof(uint32_t) * 130));
  NBX_CUDA(cudaMalloc(&s->thr_table, sizeof(T) * 264));
  NBX_CUDA(cudaMalloc(&s->hkeys, sizeof(uint64_t) * n));
  NBX_CUDA(cudaMalloc(&s->order, sizeof(uint32_t) * n));
  {
    // measured (3-D galaxy): n = 10 M double 47.9 -> 45.4 ms per step (walk 42.8 -> 39.5, the extra sort 0.6), float
    // 27.7 -> 26.6; at n = 1 M the extra sort costs what the walk gains. NBX_OCT_HILBERT=0|1 forces it (experiments).
    const char* v = getenv("NBX_OCT_HILBERT");
    s->hilbert_targets = v ? atoi(v) != 0 : e->n >= (4u << 20);
  }
  NBX_CUDA(cudaMalloc(&s->a_sorted, sizeof(vec4_t<T>) * e->n_pad));
  NBX_CUDA(cudaMemsetAsync(s->a_sorted, 0, sizeof(vec4_t<T>) * e->n_pad, e->stream));
  NBX_TRY(sorter_create(e, e->n));
  return NBX_OK;
}

template <typename T, int D>
static void destroy_impl(nbx_engine* e) {
  auto* s = st<T>(e);
  if (!s) return;
  void* bufs[] = {s->root, s->partial, s->keys, s->skeys, s->keys_lo, s->skeys_lo, s->keys_tmp, s->perm_tmp, s->perm, s->delta, s->cnt, s->blocksum,
                  s->mono, s->meta, s->rec_body, s->cell_pos, s->cells_by_depth, s->depth_count, s->depth_cursor, s->a_sorted, s->thr_table, s->hkeys, s->order};
  for (void* b : bufs)
    if (b) cudaFree(b);
  delete s;
  e->octree = nullptr;
}

// keys -> sort -> delta -> [lane order] -> scan -> records, with one-word (MAXL levels) or two-word (2*MAXL) path keys.
// Everything is enqueued; nothing here waits for the device.
template <typename T, int D>
static int build_attempt(nbx_engine* e, bool deep) {
  auto* s = st<T>(e);
  const uint32_t n = e->n;
  const vec4_t<T>* xm = static_cast<const vec4_t<T>*>(e->xm[e->cur]);
  const unsigned gb = (n + 255) / 256;
  s->deep = deep;
  {
    PhaseTimer pt(e, PH_SORT);
    if (!deep) {
      path_keys_kernel<T, D><<<gb, 256, 0, e->stream>>>(xm, n, s->root, s->keys, nullptr);
      e->launches++;
      NBX_TRY(sort_pairs(e, s->keys, n, KeyTraits<D>::BITS, s->perm, s->skeys));
      delta_kernel<D><<<(n + 1 + 255) / 256, 256, 0, e->stream>>>(s->skeys, nullptr, n, s->delta, s->cnt, &s->root->overflow);
      e->launches++;
    } else {  // LSD over two words: sort by the low word, then (stably) by the high word
      NBX_CUDA(cudaMemsetAsync(&s->root->overflow, 0, sizeof(uint32_t), e->stream));
      path_keys_kernel<T, D><<<gb, 256, 0, e->stream>>>(xm, n, s->root, s->keys, s->keys_lo);
      NBX_TRY(sort_pairs(e, s->keys_lo, n, KeyTraits<D>::BITS, s->perm_tmp, nullptr));
      gather_u64_kernel<<<gb, 256, 0, e->stream>>>(s->keys, s->perm_tmp, n, s->keys_tmp);
      NBX_TRY(sort_pairs(e, s->keys_tmp, n, KeyTraits<D>::BITS, s->perm, s->skeys, s->perm_tmp));
      gather_u64_kernel<<<gb, 256, 0, e->stream>>>(s->keys_lo, s->perm, n, s->skeys_lo);
      delta_kernel<D><<<(n + 1 + 255) / 256, 256, 0, e->stream>>>(s->skeys, s->skeys_lo, n, s->delta, s->cnt, &s->root->overflow);
      e->launches += 4;
    }
    if (s->hilbert_targets) {  // lane order of the walk (see walk_order_keys_kernel)
      static const int hb_env = [] { const char* v = getenv("NBX_OCT_HB"); return v ? atoi(v) : 0; }();
      const int hb = std::min(hb_env > 0 ? hb_env : (D == 3 ? 10 : 16), D == 3 ? 21 : 32);  // <= MAXL levels, <= 64 key bits
      walk_order_keys_kernel<D><<<gb, 256, 0, e->stream>>>(s->skeys, n, hb, s->hkeys);
      e->launches++;
      NBX_TRY(sort_pairs(e, s->hkeys, n, hb * D, s->order, nullptr));
    }
  }
  {
    PhaseTimer pt(e, PH_BUILD);
    const uint32_t count = n + 1, nb = (count + SC_TILE - 1) / SC_TILE;
    scan_reduce_kernel<<<nb, 1024, 0, e->stream>>>(s->cnt, count, s->blocksum);
    scan_blocksums_kernel<<<1, 1024, 0, e->stream>>>(s->blocksum, nb, nullptr);
    scan_apply_kernel<<<nb, 1024, 0, e->stream>>>(s->cnt, count, s->blocksum);
    NBX_CUDA(cudaMemsetAsync(s->depth_count, 0, sizeof(uint32_t) * 130, e->stream));
    emit_records_kernel<T, D><<<gb, 256, 0, e->stream>>>(s->skeys, deep ? s->skeys_lo : nullptr, s->perm, xm, n, s->delta, s->cnt,
                                                        s->cap, s->ccap, s->mono, s->meta, s->rec_body, s->cell_pos, s->root, s->depth_count);
    e->launches += 4;
  }
  return NBX_OK;
}

template <typename T, int D>
static int build_impl(nbx_engine* e) {
  auto* s = st<T>(e);
  const uint32_t n = e->n;
  const vec4_t<T>* xm = static_cast<const vec4_t<T>*>(e->xm[e->cur]);
  s->built = false;
  {
    PhaseTimer pt(e, PH_BBOX);
    bounds_partial_kernel<T, D><<<s->nblocks, 256, 0, e->stream>>>(xm, n, s->partial);
    bounds_final_kernel<T><<<1, 32, 0, e->stream>>>(s->partial, s->nblocks, s->root);
    e->launches += 2;
  }
  // MAXL levels (21 in 3-D, 32 in 2-D) separate all bodies of the usual workloads. The ONE host synchronisation of the
  // build sits after the records have been emitted and reads both verdicts at once: bit 0 = two bodies still share a
  // cell (redo with two-word keys, 2*MAXL levels), bit 1 = more internal cells than capacity. Nothing that depends on a
  // valid tree (monopoles, walk) is enqueued before the verdict, and an unusable tree is reported from here, before any
  // force kernel can walk it.
  NBX_TRY((build_attempt<T, D>(e, false)));
  uint32_t hcount[130];  // cells per depth of the build (emit_records), read together with the verdict
  auto verdict = [&](uint32_t* ovf) -> int {
    NBX_CUDA(cudaMemcpyAsync(ovf, &s->root->overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    NBX_CUDA(cudaMemcpyAsync(hcount, s->depth_count, sizeof(hcount), cudaMemcpyDeviceToHost, e->stream));
    NBX_CUDA(cudaStreamSynchronize(e->stream));
    e->d2h += sizeof(uint32_t) + sizeof(hcount);
    return NBX_OK;
  };
  uint32_t ovf = 0;
  NBX_TRY(verdict(&ovf));
  if (ovf & 1u) {
    NBX_TRY((build_attempt<T, D>(e, true)));
    NBX_TRY(verdict(&ovf));
  }
  if (ovf)
    return fail(NBX_ERR_CAPACITY,
                "octree: bodies not separated within the supported depth (coincident bodies?) or more cells than the "
                "reference's capacity max(2^D*n,1000) (src/system.h:29) allows");
  {
    PhaseTimer pt(e, PH_MONO);
    depth_offsets_kernel<<<1, 32, 0, e->stream>>>(s->depth_count, s->depth_cursor);
    const unsigned gc = (s->ccap + 255) / 256;
    group_cells_kernel<T><<<gc, 256, 0, e->stream>>>(s->cell_pos, s->root, s->meta, s->cap, s->depth_cursor, s->cells_by_depth);
    e->launches += 2;
    // one launch per populated depth, deepest first, sized by that depth's cell count; the sparse top of the tree (every
    // depth from `top` up holding <= 1024 cells) goes into one single-CTA launch
    int deepest = -1, top = -1;
    for (int d = 0; d < 128; ++d)
      if (hcount[d]) deepest = d;
    for (int d = 0; d <= deepest && hcount[d] <= 1024u; ++d) top = d;
    for (int depth = deepest; depth > top; --depth) {
      if (!hcount[depth]) continue;
      monopole_level_kernel<T, D><<<(hcount[depth] + 255) / 256, 256, 0, e->stream>>>(s->cells_by_depth, s->depth_count, uint32_t(depth), s->cap,
                                                                                     s->mono, s->meta);
      e->launches++;
    }
    if (top >= 0) {
      monopole_top_kernel<T, D><<<1, 1024, 0, e->stream>>>(s->cells_by_depth, s->depth_count, top, s->cap, s->mono, s->meta);
      e->launches++;
    }
  }
  NBX_CUDA(cudaGetLastError());
  s->built = true;
  return NBX_OK;
}

template <typename T, int D>
static int check_overflow(nbx_engine* e) {
  auto* s = st<T>(e);
  Root<T> h;
  NBX_CUDA(cudaMemcpyAsync(&h, s->root, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
  NBX_CUDA(cudaStreamSynchronize(e->stream));
  e->d2h += sizeof(h);
  if (h.overflow)
    return fail(NBX_ERR_CAPACITY,
                "octree: bodies not separated within the supported depth (coincident bodies?) or more cells than the "
                "reference's capacity max(2^D*n,1000) (src/system.h:29) allows");
  return NBX_OK;
}

// bodies per warp of the walks: at small n the walk ends with its longest warp's chain of dependent steps, and the union
// path of fewer neighbours is shorter (as for the BVH, nbx_bvh.cu walk_lanes). The octree gains less — its build is half of
// a small step. Measured (B200, step ms for 32 / 16 / 8 bodies per warp, tools/exp_oct_small.py): double n = 10 k 0.499 /
// 0.470 / 0.447, 30 k 0.659 / 0.645 / 0.629, 100 k 0.968 / 0.948 / 1.146; float 10 k 0.381 / 0.373 / 0.369, 30 k 0.441 /
// 0.437 / 0.443, 100 k 0.667 / 0.684 / 0.815. Results are bit-identical (checked there). Look-ahead sector touches of the
// records 1, 2 or 4 steps on (the walk mostly moves to p + 1) were measured with it: no gain at 10 k - 30 k, 5 - 25 % slower
// from 100 k; not kept. NBX_OCT_LANES=8|16|32 forces the width (experiments).
static uint32_t oct_walk_lanes(uint32_t targets) {
  const char* v = getenv("NBX_OCT_LANES");
  const int k   = v ? atoi(v) : 0;
  if (k == 8 || k == 16 || k == 32) return uint32_t(k);
  return targets <= 20000u ? 8u : (targets <= 40000u ? 16u : 32u);
}
int octree_walk_width(const nbx_engine* e) { return int(oct_walk_lanes(e->te - e->tb)); }

template <typename T, int D, bool COUNT>
static void launch_oct_walk(nbx_engine* e, OctreeState<T>* s, int walk, unsigned long long* stats) {
  const uint32_t nt    = e->te - e->tb;
  const uint32_t lanes = oct_walk_lanes(nt);
  const unsigned grid  = (nt + 4u * lanes - 1) / (4u * lanes);
  const uint32_t* ord  = s->hilbert_targets ? s->order : nullptr;
  if (walk == 2)
    octree_force_thr_kernel<T, D, COUNT><<<grid, 128, 0, e->stream>>>(s->mono, s->meta, s->root, s->cnt, ord, e->n, e->tb, e->te, s->thr_table,
                                                                      T(e->cfg.G), s->a_sorted, stats, lanes);
  else
    octree_force_kernel<T, D, COUNT><<<grid, 128, 0, e->stream>>>(s->mono, s->meta, s->root, s->cnt, ord, e->n, e->tb, e->te, s->thr_table,
                                                                  T(e->cfg.G), s->a_sorted, stats, lanes);
}

template <typename T, int D>
static int force_impl(nbx_engine* e) {
  auto* s = st<T>(e);
  if (!s->built) return fail(NBX_ERR_STATE, "octree compute_force before build");
  const uint32_t nt = e->te - e->tb;
  PhaseTimer pt(e, PH_TRAVERSE);
  if (nt) {
    // double: threshold form (measured 42.7 vs 51.9 ms at n = 10 M); float: the sqrt form stays branch-free and is faster
    // (26.0 vs 29.8 ms). NBX_OCT_WALK=1|2 forces one of them (experiments).
    static const int forced = [] { const char* v = getenv("NBX_OCT_WALK"); return v ? atoi(v) : 0; }();
    const int walk = forced ? forced : (sizeof(T) == 8 ? 2 : 1);
    threshold_table_kernel<T><<<1, 160, 0, e->stream>>>(s->root, T(e->cfg.theta), s->thr_table);
    launch_oct_walk<T, D, false>(e, s, walk, nullptr);
    e->launches += 2;
  }
  if (e->cfg.world_size > 1) NBX_TRY(comm_allgather(e, s->a_sorted));
  unsort_kernel<T><<<(e->n + 255) / 256, 256, 0, e->stream>>>(s->perm, s->hilbert_targets ? s->order : nullptr, s->a_sorted, e->n,
                                                              static_cast<vec4_t<T>*>(e->a));
  e->launches++;
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

template <typename T, int D>
static int get_root_impl(nbx_engine* e, void* side, void* root_x, uint64_t* nodes_used) {
  auto* s = st<T>(e);
  if (!s->built) return fail(NBX_ERR_STATE, "no octree build has run yet");
  NBX_TRY((check_overflow<T, D>(e)));
  Root<T> h;
  NBX_CUDA(cudaMemcpy(&h, s->root, sizeof(h), cudaMemcpyDeviceToHost));
  if (side) *static_cast<T*>(side) = h.side;
  if (root_x)
    for (int k = 0; k < D; ++k) static_cast<T*>(root_x)[k] = h.x;
  if (nodes_used) *nodes_used = 1 + (uint64_t(1) << D) * h.cells;  // what next_free_child_group reads (octree.h:152)
  return NBX_OK;
}

template <typename T, int D>
static int get_canonical_impl(nbx_engine* e, uint64_t* count, uint32_t* depth, uint64_t* path, uint32_t* kind, void* monopole) {
  auto* s = st<T>(e);
  if (!s->built) return fail(NBX_ERR_STATE, "no octree build has run yet");
  NBX_TRY((check_overflow<T, D>(e)));
  Root<T> h;
  NBX_CUDA(cudaMemcpy(&h, s->root, sizeof(h), cudaMemcpyDeviceToHost));
  const uint32_t nrec = e->n + h.cells;
  if (count) *count = nrec;
  if (!depth && !path && !kind && !monopole) return NBX_OK;
  if (!depth || !path || !kind || !monopole) return fail(NBX_ERR_INVALID, "octree_get_canonical: pass all arrays or none");
  uint32_t *dd = nullptr, *dk = nullptr;
  uint64_t* dp = nullptr;
  T* dm        = nullptr;
  NBX_CUDA(cudaMalloc(&dd, sizeof(uint32_t) * nrec));
  NBX_CUDA(cudaMalloc(&dk, sizeof(uint32_t) * nrec));
  NBX_CUDA(cudaMalloc(&dp, sizeof(uint64_t) * nrec));
  NBX_CUDA(cudaMalloc(&dm, sizeof(T) * size_t(nrec) * (D + 1)));
  export_canonical_kernel<T, D><<<(nrec + 255) / 256, 256, 0, e->stream>>>(s->mono, s->meta, s->rec_body, s->skeys, nrec, dd, dp, dk, dm);
  e->launches++;
  cudaError_t err = cudaMemcpyAsync(depth, dd, sizeof(uint32_t) * nrec, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(kind, dk, sizeof(uint32_t) * nrec, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(path, dp, sizeof(uint64_t) * nrec, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(monopole, dm, sizeof(T) * size_t(nrec) * (D + 1), cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  e->d2h += size_t(nrec) * (16 + sizeof(T) * (D + 1));
  cudaFree(dd); cudaFree(dk); cudaFree(dp); cudaFree(dm);
  if (err != cudaSuccess) return fail(NBX_ERR_CUDA, cudaGetErrorString(err));
  return NBX_OK;
}

template <typename T, int D>
static int stats_impl(nbx_engine* e, unsigned long long* dev_stats) {
  auto* s = st<T>(e);
  if (!s->built) return fail(NBX_ERR_STATE, "no octree build has run yet");
  const uint32_t nt = e->te - e->tb;
  if (nt) {  // the counting twin of the walk force_impl launches
    threshold_table_kernel<T><<<1, 160, 0, e->stream>>>(s->root, T(e->cfg.theta), s->thr_table);
    launch_oct_walk<T, D, true>(e, s, sizeof(T) == 8 ? 2 : 1, dev_stats);
    e->launches += 2;
  }
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

#define OCT_DISPATCH(e, fn, ...)                                                          \
  ((e)->prec == 4 ? ((e)->dim == 2 ? fn<float, 2>(__VA_ARGS__) : fn<float, 3>(__VA_ARGS__)) \
                  : ((e)->dim == 2 ? fn<double, 2>(__VA_ARGS__) : fn<double, 3>(__VA_ARGS__)))

int octree_create(nbx_engine* e) { return OCT_DISPATCH(e, create_impl, e); }
void octree_destroy(nbx_engine* e) {
  if (!e->octree) return;
  OCT_DISPATCH(e, destroy_impl, e);
}
int octree_build(nbx_engine* e) { return OCT_DISPATCH(e, build_impl, e); }
int octree_compute_force(nbx_engine* e) { return OCT_DISPATCH(e, force_impl, e); }
int octree_check(nbx_engine* e) { return OCT_DISPATCH(e, check_overflow, e); }
int octree_stats(nbx_engine* e, unsigned long long* dev_stats) { return OCT_DISPATCH(e, stats_impl, e, dev_stats); }
int octree_get_root(nbx_engine* e, void* side, void* root_x, uint64_t* nodes_used) {
  return OCT_DISPATCH(e, get_root_impl, e, side, root_x, nodes_used);
}
int octree_get_canonical(nbx_engine* e, uint64_t* count, uint32_t* depth, uint64_t* path, uint32_t* kind, void* monopole) {
  return OCT_DISPATCH(e, get_canonical_impl, e, count, depth, path, kind, monopole);
}

}  // namespace nbx
