// nbx_sort.cu — stable LSD radix sort of (u64 key, u32 index) pairs, hand-written for sm_100a (no CUB/Thrust).
//
// Replaces the reference's std::sort(par_unseq, pair<key,idx>) (src/bvh.h:62-69). 8 bits per pass; each pass is
//   digit_histogram  : per-tile digit counts            -> hist[digit][tile]
//   row_scan         : exclusive scan of every digit row (+ digit totals)
//   scatter          : warp-match ranking (stable) + global offsets; the tile is first staged in shared memory in sorted
//                      order so that the write-out is coalesced per digit run
// A tile is RS_THREADS x RS_ITEMS consecutive elements laid out warp-striped, so loads are fully coalesced and the
// element order inside a tile is (warp, iteration, lane) — ranks are assigned in exactly that order => STABLE, so ties
// keep their current index order (the reference's unstable std::sort leaves ties undefined).
// HBM traffic per pass: 8 B (histogram) + 12 B read + 12 B write per element.
#include "nbx_internal.cuh"

namespace nbx {

namespace {

constexpr int RS_THREADS = 512;
constexpr int RS_ITEMS   = 8;
constexpr int RS_TILE    = RS_THREADS * RS_ITEMS;  // 4096 elements
constexpr int RS_WARPS   = RS_THREADS / 32;
constexpr int RS_RADIX   = 256;

struct Sorter {
  uint32_t capacity = 0;
  uint32_t ntiles   = 0;
  uint64_t* keys[2] = {nullptr, nullptr};
  uint32_t* vals[2] = {nullptr, nullptr};
  uint32_t* hist    = nullptr;  // [RS_RADIX][ntiles]
  uint32_t* totals  = nullptr;  // [RS_RADIX]
};

__global__ void __launch_bounds__(RS_THREADS) digit_histogram_kernel(const uint64_t* __restrict__ keys, uint32_t n,
                                                                     int shift, uint32_t* __restrict__ hist,
                                                                     uint32_t ntiles) {
  __shared__ uint32_t h[RS_RADIX];
  for (int d = threadIdx.x; d < RS_RADIX; d += RS_THREADS) h[d] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    uint32_t idx = base + (threadIdx.x >> 5) * (32 * RS_ITEMS) + k * 32 + (threadIdx.x & 31);
    if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 0xff], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RS_RADIX; d += RS_THREADS) hist[size_t(d) * ntiles + blockIdx.x] = h[d];
}

// one CTA per digit: exclusive scan of hist[d][0..ntiles) in place, total -> totals[d]
__global__ void __launch_bounds__(1024) row_scan_kernel(uint32_t* hist, uint32_t ntiles, uint32_t* totals) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  uint32_t* row = hist + size_t(blockIdx.x) * ntiles;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < ntiles; base += 1024) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < ntiles ? row[i] : 0;
    uint32_t s = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, s, off);
      if (lane >= off) s += t;
    }
    if (lane == 31) warp_sums[warp] = s;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, w, off);
        if (lane >= off) w += t;
      }
      warp_sums[lane] = w;  // inclusive
    }
    __syncthreads();
    uint32_t prefix = carry + (warp ? warp_sums[warp - 1] : 0) + (s - v);
    if (i < ntiles) row[i] = prefix;
    __syncthreads();
    if (threadIdx.x == 1023) carry = prefix + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// dynamic shared memory of scatter_kernel: per-warp digit counters + the tile staged in sorted order
constexpr size_t RS_SMEM = sizeof(uint32_t) * RS_WARPS * RS_RADIX + sizeof(uint64_t) * RS_TILE + sizeof(uint32_t) * RS_TILE;

__global__ void __launch_bounds__(RS_THREADS) scatter_kernel(const uint64_t* __restrict__ keys_in,
                                                             const uint32_t* __restrict__ vals_in,  // NULL => iota
                                                             uint64_t* __restrict__ keys_out,
                                                             uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                             const uint32_t* __restrict__ hist, uint32_t ntiles,
                                                             const uint32_t* __restrict__ totals) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  uint32_t(*warp_count)[RS_RADIX] = reinterpret_cast<uint32_t(*)[RS_RADIX]>(rs_smem);
  uint64_t* skey                  = reinterpret_cast<uint64_t*>(rs_smem + sizeof(uint32_t) * RS_WARPS * RS_RADIX);
  uint32_t* sval                  = reinterpret_cast<uint32_t*>(skey + RS_TILE);
  __shared__ uint32_t gbase[RS_RADIX];   // global position of this tile's first element of digit d
  __shared__ uint32_t dstart[RS_RADIX];  // position of digit d inside the sorted tile
  __shared__ uint32_t scan_tmp[RS_RADIX / 32];
  __shared__ uint32_t scan_tmp2[RS_RADIX / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = threadIdx.x; q < RS_WARPS * RS_RADIX; q += RS_THREADS) (&warp_count[0][0])[q] = 0;

  // exclusive scan of the 256 digit totals (every CTA recomputes it: 256 values)
  uint32_t tot = 0, inc = 0;
  if (threadIdx.x < RS_RADIX) {
    tot = totals[threadIdx.x];
    inc = tot;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += t;
    }
    if (lane == 31) scan_tmp[warp] = inc;
  }
  __syncthreads();
  if (threadIdx.x < RS_RADIX) {
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += scan_tmp[w];
    gbase[threadIdx.x] = before + inc - tot + hist[size_t(threadIdx.x) * ntiles + blockIdx.x];
  }

  uint64_t key[RS_ITEMS];
  uint32_t val[RS_ITEMS], rank[RS_ITEMS];
  const uint32_t tile0 = blockIdx.x * RS_TILE;
  const uint32_t base  = tile0 + warp * (32 * RS_ITEMS) + lane;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    uint32_t idx = base + k * 32;
    key[k]       = idx < n ? keys_in[idx] : ~0ull;
    val[k]       = idx < n ? (vals_in ? vals_in[idx] : idx) : 0xffffffffu;
  }
  __syncthreads();  // warp_count zeroed, gbase ready
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    uint32_t d     = uint32_t(key[k] >> shift) & 0xff;
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t lt    = peers & ((1u << lane) - 1);
    uint32_t cnt   = warp_count[warp][d];  // every peer reads the same value before the leader bumps it
    __syncwarp();
    if (lt == 0) warp_count[warp][d] = cnt + __popc(peers);
    __syncwarp();
    rank[k] = cnt + __popc(lt);
  }
  __syncthreads();
  // per digit: exclusive prefix over the warps of this CTA, then over the digits (position inside the sorted tile)
  uint32_t dcount = 0, dinc = 0;
  if (threadIdx.x < RS_RADIX) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      uint32_t c                 = warp_count[w][threadIdx.x];
      warp_count[w][threadIdx.x] = run;
      run += c;
    }
    dcount = run;
    dinc   = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, dinc, off);
      if (lane >= off) dinc += t;
    }
    if (lane == 31) scan_tmp2[warp] = dinc;
  }
  __syncthreads();
  if (threadIdx.x < RS_RADIX) {
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += scan_tmp2[w];
    dstart[threadIdx.x] = before + dinc - dcount;
  }
  __syncthreads();
  // stage the tile in sorted order (stable: (warp, iteration, lane) order inside every digit)
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    uint32_t d  = uint32_t(key[k] >> shift) & 0xff;
    uint32_t lp = dstart[d] + warp_count[warp][d] + rank[k];
    skey[lp]    = key[k];
    sval[lp]    = val[k];
  }
  __syncthreads();
  // coalesced write-out: consecutive threads -> consecutive addresses inside every digit run. Elements past n carry the
  // all-ones key, sort to the very end of the tile and are simply not written.
  const uint32_t count = n - tile0 < uint32_t(RS_TILE) ? n - tile0 : uint32_t(RS_TILE);
  for (uint32_t i = threadIdx.x; i < count; i += RS_THREADS) {
    const uint64_t kk = skey[i];
    const uint32_t d  = uint32_t(kk >> shift) & 0xff;
    const uint32_t gp = gbase[d] + (i - dstart[d]);
    keys_out[gp]      = kk;
    vals_out[gp]      = sval[i];
  }
}

}  // namespace

int sorter_create(nbx_engine* e, uint32_t n) {
  if (e->sorter) return NBX_OK;
  Sorter* s   = new Sorter();
  e->sorter   = s;
  s->capacity = n;
  s->ntiles   = (n + RS_TILE - 1) / RS_TILE;
  for (int k = 0; k < 2; ++k) {
    NBX_CUDA(cudaMalloc(&s->keys[k], sizeof(uint64_t) * size_t(n)));
    NBX_CUDA(cudaMalloc(&s->vals[k], sizeof(uint32_t) * size_t(n)));
  }
  NBX_CUDA(cudaMalloc(&s->hist, sizeof(uint32_t) * size_t(RS_RADIX) * s->ntiles));
  NBX_CUDA(cudaMalloc(&s->totals, sizeof(uint32_t) * RS_RADIX));
  NBX_TRY(ensure_dynamic_smem(e, scatter_kernel, RS_SMEM));
  return NBX_OK;
}

void sorter_destroy(nbx_engine* e) {
  Sorter* s = static_cast<Sorter*>(e->sorter);
  if (!s) return;
  for (int k = 0; k < 2; ++k) {
    if (s->keys[k]) cudaFree(s->keys[k]);
    if (s->vals[k]) cudaFree(s->vals[k]);
  }
  if (s->hist) cudaFree(s->hist);
  if (s->totals) cudaFree(s->totals);
  delete s;
  e->sorter = nullptr;
}

int sort_pairs(nbx_engine* e, const uint64_t* keys_in, uint32_t n, int key_bits, uint32_t* perm_out,
               uint64_t* keys_sorted_out, const uint32_t* vals_in) {
  NBX_TRY(sorter_create(e, n));
  Sorter* s = static_cast<Sorter*>(e->sorter);
  if (n > s->capacity) return fail(NBX_ERR_INVALID, "sort_pairs: n exceeds sorter capacity");
  const uint32_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  int passes            = (key_bits + 7) / 8;
  if (passes < 1) passes = 1;
  if (passes > 8) passes = 8;
  const uint64_t* kin = keys_in;
  const uint32_t* vin = vals_in;  // nullptr => iota
  for (int p = 0; p < passes; ++p) {
    const bool last = p == passes - 1;
    uint64_t* kout  = last && keys_sorted_out ? keys_sorted_out : s->keys[p & 1];
    uint32_t* vout  = last ? perm_out : s->vals[p & 1];
    digit_histogram_kernel<<<ntiles, RS_THREADS, 0, e->stream>>>(kin, n, 8 * p, s->hist, ntiles);
    row_scan_kernel<<<RS_RADIX, 1024, 0, e->stream>>>(s->hist, ntiles, s->totals);
    scatter_kernel<<<ntiles, RS_THREADS, RS_SMEM, e->stream>>>(kin, vin, kout, vout, n, 8 * p, s->hist, ntiles, s->totals);
    e->launches += 3;
    kin = kout;
    vin = vout;
  }
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

}  // namespace nbx
