// nbx_sort.cu — stable LSD radix sort of (u64 key, u32 index) pairs, hand-written for sm_100a (no CUB/Thrust).
//
// Replaces the reference's std::sort(par_unseq, pair<key,idx>) (src/bvh.h:62-69). 8 bits per pass, ONE sweep over the
// data per pass ("onesweep": Adinets & Merrill, "Onesweep: a faster least significant digit radix sort for GPUs", 2022):
//   digit_histogram_kernel : one read of the keys (8 B/element) gives the digit totals of ALL passes at once; the same
//                            kernel clears the tile-status words of the sweeps
//   digit_scan_kernel      : exclusive scan of each pass's 256 totals -> first output position of every digit
//   onesweep_kernel        : per pass; a CTA takes the next tile (atomic ticket => tiles start in index order), ranks its
//                            elements per digit (warp-match ranking), publishes its per-digit counts, and obtains the
//                            number of equal-digit elements in all EARLIER tiles by decoupled look-back over the status
//                            words of its predecessors (a 32-bit word = 2 flag bits + 30-bit count, so no fence is
//                            needed); the tile is staged in shared memory in sorted order and written out coalesced per
//                            digit run.
// A tile is OS_THREADS x ITEMS consecutive elements laid out warp-striped, so loads are fully coalesced and the element
// order inside a tile is (warp, iteration, lane) = index order; ranks are assigned in exactly that order and tiles are
// prefix-summed in index order => STABLE: ties keep their current index order (the reference's unstable std::sort leaves
// ties undefined).
// HBM traffic: 8 B/element once + per pass 12 B read + 12 B written (the last pass skips the keys when nobody wants them):
// 8 + 24 * passes bytes per element, 10 launches for 64-bit keys (3 kernels x 8 passes = 24 launches and 32 B per
// element and pass before).
#include "nbx_internal.cuh"

namespace nbx {

namespace {

#ifndef NBX_OS_THREADS
#define NBX_OS_THREADS 512
#endif
#ifndef NBX_OS_ITEMS
#define NBX_OS_ITEMS 8
#endif
#ifndef NBX_OS_MINB
#define NBX_OS_MINB 1
#endif
constexpr int OS_THREADS = NBX_OS_THREADS;
constexpr int OS_WARPS   = OS_THREADS / 32;
constexpr int OS_RADIX   = 256;
constexpr int OS_ITEMS   = NBX_OS_ITEMS;
constexpr int OS_TILE    = OS_THREADS * OS_ITEMS;  // 4096 elements
constexpr int OS_MAXPASS = 8;

constexpr uint32_t ST_LOCAL = 1u << 30, ST_INCL = 2u << 30, ST_FLAGS = 3u << 30, ST_COUNT = ~ST_FLAGS;

// what a sweep partitions by: an 8-bit radix digit of the key, or — for the multi-GPU sharded sort — the rank that owns the
// key's range: the number of splitters <= key (equal keys always share an owner, so the partition composes with the
// per-rank sorts into the same stable order as one global sort)
struct RadixDigit {
  int shift;
  __device__ __forceinline__ uint32_t operator()(uint64_t k) const { return uint32_t(k >> shift) & 0xff; }
};
constexpr int OS_MAXSPLIT = 15;
struct OwnerDigit {
  uint64_t split[OS_MAXSPLIT];  // ascending; unused entries = ~0ull are never <= a real key... unless the key is ~0ull:
  int nsplit;                   // hence the explicit count
  __device__ __forceinline__ uint32_t operator()(uint64_t k) const {
    uint32_t d = 0;
#pragma unroll
    for (int q = 0; q < OS_MAXSPLIT; ++q) d += (q < nsplit) & (split[q] <= k);
    return d;
  }
};

struct Sorter {
  uint32_t capacity = 0;
  uint32_t ntiles   = 0;
  uint64_t* keys[2] = {nullptr, nullptr};
  uint32_t* vals[2] = {nullptr, nullptr};
  uint32_t* ghist   = nullptr;  // [OS_MAXPASS][OS_RADIX] digit totals -> exclusive offsets; followed by [OS_MAXPASS] tickets
  uint32_t* status  = nullptr;  // [OS_MAXPASS][ntiles][OS_RADIX]
  uint64_t* samples = nullptr;  // sharded sort: [2][OS_SAMPLES] sampled keys / sorted
  uint32_t* sample_perm = nullptr;
};
constexpr uint32_t OS_SAMPLES = 1u << 16;

// digit totals of every pass in one read of the keys; also clears the status words of the sweeps that follow
__global__ void __launch_bounds__(OS_THREADS) digit_histogram_kernel(const uint64_t* __restrict__ keys, uint32_t n, int passes,
                                                                     uint32_t* __restrict__ ghist, uint32_t* __restrict__ status,
                                                                     size_t status_words) {
  __shared__ uint32_t h[OS_MAXPASS][OS_RADIX];
  for (int q = threadIdx.x; q < OS_MAXPASS * OS_RADIX; q += OS_THREADS) (&h[0][0])[q] = 0;
  for (size_t q = size_t(blockIdx.x) * OS_THREADS + threadIdx.x; q < status_words / 4; q += size_t(gridDim.x) * OS_THREADS)
    reinterpret_cast<uint4*>(status)[q] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (uint32_t i = blockIdx.x * OS_THREADS + threadIdx.x; i < n; i += gridDim.x * OS_THREADS) {
    const uint64_t k = keys[i];
#pragma unroll
    for (int p = 0; p < OS_MAXPASS; ++p)
      if (p < passes) atomicAdd(&h[p][(k >> (8 * p)) & 0xff], 1u);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < passes * OS_RADIX; q += OS_THREADS) {
    const uint32_t c = (&h[0][0])[q];
    if (c) atomicAdd(&ghist[q], c);
  }
}

// sharded sort: elements per owner (one row of ghist), and the clearing of that sweep's status words
__global__ void __launch_bounds__(OS_THREADS) owner_histogram_kernel(const uint64_t* __restrict__ keys, uint32_t n, const OwnerDigit digit,
                                                                     uint32_t* __restrict__ ghist, uint32_t* __restrict__ status,
                                                                     size_t status_words) {
  __shared__ uint32_t h[OS_MAXSPLIT + 1];
  if (threadIdx.x <= OS_MAXSPLIT) h[threadIdx.x] = 0;
  for (size_t q = size_t(blockIdx.x) * OS_THREADS + threadIdx.x; q < status_words / 4; q += size_t(gridDim.x) * OS_THREADS)
    reinterpret_cast<uint4*>(status)[q] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  uint32_t mine[OS_MAXSPLIT + 1];
#pragma unroll
  for (int q = 0; q <= OS_MAXSPLIT; ++q) mine[q] = 0;
  for (uint32_t i = blockIdx.x * OS_THREADS + threadIdx.x; i < n; i += gridDim.x * OS_THREADS) {
    const uint32_t d = digit(keys[i]);
#pragma unroll
    for (int q = 0; q <= OS_MAXSPLIT; ++q) mine[q] += d == uint32_t(q);
  }
#pragma unroll
  for (int q = 0; q <= OS_MAXSPLIT; ++q) {
    uint32_t v = mine[q];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&h[q], v);
  }
  __syncthreads();
  if (threadIdx.x <= OS_MAXSPLIT && h[threadIdx.x]) atomicAdd(&ghist[threadIdx.x], h[threadIdx.x]);
}

// evenly spaced sample of the keys (the same on every rank)
__global__ void __launch_bounds__(256) sample_keys_kernel(const uint64_t* __restrict__ keys, uint32_t n, uint32_t nsamples,
                                                          uint64_t* __restrict__ samples) {
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i < nsamples) samples[i] = keys[uint64_t(i) * n / nsamples];
}

// one CTA per pass: exclusive scan of its 256 digit totals, in place
__global__ void __launch_bounds__(OS_RADIX) digit_scan_kernel(uint32_t* ghist) {
  __shared__ uint32_t ws[OS_RADIX / 32];
  uint32_t* row = ghist + blockIdx.x * OS_RADIX;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t v = row[threadIdx.x];
  uint32_t s = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, s, off);
    if (lane >= off) s += t;
  }
  if (lane == 31) ws[warp] = s;
  __syncthreads();
  uint32_t before = 0;
  for (int w = 0; w < warp; ++w) before += ws[w];
  row[threadIdx.x] = before + s - v;
}

// lanes of the warp holding the same 8-bit digit. MATCH.ANY does this in one instruction, but on B200 it sustains only about
// one warp instruction per 114 cycles per scheduler (measured with %globaltimer inside this kernel: 3.7 us of a 9.5 us tile for
// 8 matches per thread); eight ballots — one per digit bit, each ANDed in plain or complemented — cost about 32 ALU
// instructions per item and are faster.
__device__ __forceinline__ uint32_t digit_peers(uint32_t d) {
#ifdef NBX_OS_USE_MATCH
  return __match_any_sync(0xffffffffu, d);
#else
  // four instructions per bit (test -> predicate, ballot, select the complement mask, combine); written in PTX because the
  // compiler otherwise re-derives each bit from the 64-bit key shift and spends 5-6
  uint32_t peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; ++b)
    asm("{\n"
        ".reg .pred p;\n"
        ".reg .b32 t, bal, m;\n"
        "and.b32 t, %1, %2;\n"
        "setp.ne.u32 p, t, 0;\n"
        "vote.sync.ballot.b32 bal, p, 0xffffffff;\n"
        "selp.b32 m, 0, 0xffffffff, p;\n"
        "xor.b32 bal, bal, m;\n"  // lanes whose bit b equals mine
        "and.b32 %0, %0, bal;\n"
        "}\n"
        : "+r"(peers)
        : "r"(d), "r"(1u << b));
  return peers;
#endif
}

// dynamic shared memory of onesweep_kernel: per-warp digit counters + the tile staged in sorted order
constexpr size_t OS_SMEM = sizeof(uint32_t) * OS_WARPS * OS_RADIX + sizeof(uint64_t) * OS_TILE + sizeof(uint32_t) * OS_TILE;

template <typename Digit>
__global__ void __launch_bounds__(OS_THREADS, NBX_OS_MINB) onesweep_kernel(const uint64_t* __restrict__ keys_in,
                                                              const uint32_t* __restrict__ vals_in,  // NULL => iota
                                                              uint64_t* __restrict__ keys_out,       // NULL => not wanted
                                                              uint32_t* __restrict__ vals_out, uint32_t n, const Digit digit,
                                                              const uint32_t* __restrict__ digit_offset, uint32_t* status,
                                                              uint32_t* ticket) {
  extern __shared__ __align__(16) unsigned char os_smem[];
  uint32_t(*warp_count)[OS_RADIX] = reinterpret_cast<uint32_t(*)[OS_RADIX]>(os_smem);
  uint64_t* skey                  = reinterpret_cast<uint64_t*>(os_smem + sizeof(uint32_t) * OS_WARPS * OS_RADIX);
  uint32_t* sval                  = reinterpret_cast<uint32_t*>(skey + OS_TILE);
  __shared__ uint32_t goff[OS_RADIX];    // global position of this tile's first element of digit d, minus dstart[d]
  __shared__ uint32_t dstart[OS_RADIX];  // position of digit d inside the sorted tile
  __shared__ uint32_t scan_tmp[OS_RADIX / 32];
  __shared__ uint32_t s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);  // tiles are handed out in launch order: every predecessor runs
  for (int q = threadIdx.x; q < OS_WARPS * OS_RADIX; q += OS_THREADS) (&warp_count[0][0])[q] = 0;
  __syncthreads();
  const uint32_t tile  = s_tile;
  const uint32_t tile0 = tile * OS_TILE;

  uint64_t key[OS_ITEMS];
  uint32_t val[OS_ITEMS], rank[OS_ITEMS];
  const uint32_t base = tile0 + warp * (32 * OS_ITEMS) + lane;
#pragma unroll
  for (int k = 0; k < OS_ITEMS; ++k) {
    const uint32_t idx = base + k * 32;
    key[k]             = idx < n ? keys_in[idx] : ~0ull;
    val[k]             = idx < n ? (vals_in ? vals_in[idx] : idx) : 0xffffffffu;
  }
  // warp-match ranking: all lanes of a warp holding digit d in item k form `peers`; the lowest of them bumps the warp's
  // counter of d by the group size (one shared-memory atomic, distinct addresses within the instruction) and hands the
  // old value to its peers with a shuffle. The 8 matches are independent and issue back to back; the atomics of one
  // warp execute in program order, so ranks follow (item, lane) order.
  uint32_t peers[OS_ITEMS];
#pragma unroll
#ifndef NBX_OS_NMATCH
#define NBX_OS_NMATCH 0
#endif
  for (int k = 0; k < OS_ITEMS; ++k)  // the first NBX_OS_NMATCH items on the MATCH unit, the others on ballots
    peers[k] = k < NBX_OS_NMATCH ? __match_any_sync(0xffffffffu, digit(key[k])) : digit_peers(digit(key[k]));
#pragma unroll
  for (int k = 0; k < OS_ITEMS; ++k) {
    const uint32_t d      = digit(key[k]);
    const uint32_t lt     = peers[k] & ((1u << lane) - 1);
    const int leader      = __ffs(peers[k]) - 1;
    uint32_t old          = 0;
    if (lt == 0) old = atomicAdd(&warp_count[warp][d], uint32_t(__popc(peers[k])));
    rank[k] = __shfl_sync(0xffffffffu, old, leader) + __popc(lt);
  }
  __syncthreads();
  // per digit (thread d): exclusive prefix over the warps of this CTA, publish the tile's count, position of the digit
  // inside the sorted tile
  uint32_t dcount = 0, dinc = 0;
  uint32_t* my_status = status + (size_t(tile) * OS_RADIX + threadIdx.x);
  if (threadIdx.x < OS_RADIX) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) {
      const uint32_t c           = warp_count[w][threadIdx.x];
      warp_count[w][threadIdx.x] = run;
      run += c;
    }
    dcount = run;
    *reinterpret_cast<volatile uint32_t*>(my_status) = (tile == 0 ? ST_INCL : ST_LOCAL) | dcount;
    dinc = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, dinc, off);
      if (lane >= off) dinc += t;
    }
    if (lane == 31) scan_tmp[warp] = dinc;
  }
  __syncthreads();
  uint32_t my_dstart = 0;
  if (threadIdx.x < OS_RADIX) {
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += scan_tmp[w];
    my_dstart           = before + dinc - dcount;
    dstart[threadIdx.x] = my_dstart;
  }
  __syncthreads();
  // decoupled look-back (digit threads): elements of digit d in all earlier tiles. It overlaps the staging done by the
  // other warps; a predecessor that has not published yet is simply polled again.
  if (threadIdx.x < OS_RADIX) {
    // LB predecessors are read at once (independent loads in flight), then consumed nearest first; one that has not
    // published yet is polled again. Rows before tile 0 count as "inclusive 0".
#ifndef NBX_OS_LB
#define NBX_OS_LB 4
#endif
    constexpr int LB = NBX_OS_LB;
    uint32_t excl = 0;
    bool done     = tile == 0;
    for (int64_t t = int64_t(tile) - 1; !done; t -= LB) {
      uint32_t v[LB];
#pragma unroll
      for (int q = 0; q < LB; ++q)
        v[q] = t - q >= 0 ? *reinterpret_cast<const volatile uint32_t*>(status + (size_t(t - q) * OS_RADIX + threadIdx.x)) : ST_INCL;
#pragma unroll
      for (int q = 0; q < LB; ++q) {
        if (!done) {
          while ((v[q] & ST_FLAGS) == 0) {
#ifdef NBX_OS_SLEEP
            __nanosleep(NBX_OS_SLEEP);
#endif
            v[q] = *reinterpret_cast<const volatile uint32_t*>(status + (size_t(t - q) * OS_RADIX + threadIdx.x));
          }
          excl += v[q] & ST_COUNT;
          done = (v[q] & ST_INCL) != 0;
        }
      }
    }
    if (tile != 0) *reinterpret_cast<volatile uint32_t*>(my_status) = ST_INCL | (excl + dcount);
    goff[threadIdx.x] = digit_offset[threadIdx.x] + excl - my_dstart;
  }
  // stage the tile in sorted order (stable: (warp, iteration, lane) order inside every digit)
#pragma unroll
  for (int k = 0; k < OS_ITEMS; ++k) {
    const uint32_t d  = digit(key[k]);
    const uint32_t lp = dstart[d] + warp_count[warp][d] + rank[k];
    skey[lp]          = key[k];
    sval[lp]          = val[k];
  }
  __syncthreads();
  // coalesced write-out: consecutive threads -> consecutive addresses inside every digit run. Elements past n carry the
  // all-ones key, sort to the very end of the (last) tile and are simply not written.
  const uint32_t count = n - tile0 < uint32_t(OS_TILE) ? n - tile0 : uint32_t(OS_TILE);
  for (uint32_t i = threadIdx.x; i < count; i += OS_THREADS) {
    const uint64_t kk = skey[i];
    const uint32_t gp = goff[digit(kk)] + i;
    if (keys_out) keys_out[gp] = kk;
    vals_out[gp] = sval[i];
  }
}

}  // namespace

int sorter_create(nbx_engine* e, uint32_t n) {
  if (e->sorter) return NBX_OK;
  if (n >= (1u << 30)) return fail(NBX_ERR_INVALID, "sort: n must be below 2^30 (30-bit tile status counts)");
  Sorter* s   = new Sorter();
  e->sorter   = s;
  s->capacity = n;
  s->ntiles   = (n + OS_TILE - 1) / OS_TILE;
  for (int k = 0; k < 2; ++k) {
    NBX_CUDA(cudaMalloc(&s->keys[k], sizeof(uint64_t) * size_t(n)));
    NBX_CUDA(cudaMalloc(&s->vals[k], sizeof(uint32_t) * size_t(n)));
  }
  NBX_CUDA(cudaMalloc(&s->ghist, sizeof(uint32_t) * (OS_MAXPASS * OS_RADIX + OS_MAXPASS)));
  NBX_CUDA(cudaMalloc(&s->status, sizeof(uint32_t) * size_t(OS_MAXPASS) * s->ntiles * OS_RADIX));
  if (e->cfg.world_size > 1) {
    NBX_CUDA(cudaMalloc(&s->samples, sizeof(uint64_t) * 2 * OS_SAMPLES));
    NBX_CUDA(cudaMalloc(&s->sample_perm, sizeof(uint32_t) * OS_SAMPLES));
  }
  NBX_TRY(ensure_dynamic_smem(e, onesweep_kernel<RadixDigit>, OS_SMEM));
  NBX_TRY(ensure_dynamic_smem(e, onesweep_kernel<OwnerDigit>, OS_SMEM));
  return NBX_OK;
}

void sorter_destroy(nbx_engine* e) {
  Sorter* s = static_cast<Sorter*>(e->sorter);
  if (!s) return;
  for (int k = 0; k < 2; ++k) {
    if (s->keys[k]) cudaFree(s->keys[k]);
    if (s->vals[k]) cudaFree(s->vals[k]);
  }
  if (s->ghist) cudaFree(s->ghist);
  if (s->status) cudaFree(s->status);
  if (s->samples) cudaFree(s->samples);
  if (s->sample_perm) cudaFree(s->sample_perm);
  delete s;
  e->sorter = nullptr;
}

int sort_pairs(nbx_engine* e, const uint64_t* keys_in, uint32_t n, int key_bits, uint32_t* perm_out,
               uint64_t* keys_sorted_out, const uint32_t* vals_in) {
  NBX_TRY(sorter_create(e, n));
  Sorter* s = static_cast<Sorter*>(e->sorter);
  if (n > s->capacity) return fail(NBX_ERR_INVALID, "sort_pairs: n exceeds sorter capacity");
  const uint32_t ntiles = (n + OS_TILE - 1) / OS_TILE;
  int passes            = (key_bits + 7) / 8;
  if (passes < 1) passes = 1;
  if (passes > OS_MAXPASS) passes = OS_MAXPASS;
  uint32_t* tickets = s->ghist + OS_MAXPASS * OS_RADIX;
  NBX_CUDA(cudaMemsetAsync(s->ghist, 0, sizeof(uint32_t) * (OS_MAXPASS * OS_RADIX + OS_MAXPASS), e->stream));
  const size_t status_words = size_t(passes) * ntiles * OS_RADIX;
  const unsigned hgrid      = std::min<unsigned>(ntiles, unsigned(e->sm_count) * 4);
  digit_histogram_kernel<<<hgrid, OS_THREADS, 0, e->stream>>>(keys_in, n, passes, s->ghist, s->status, status_words);
  digit_scan_kernel<<<passes, OS_RADIX, 0, e->stream>>>(s->ghist);
  e->launches += 2;
  const uint64_t* kin = keys_in;
  const uint32_t* vin = vals_in;  // nullptr => iota
  for (int p = 0; p < passes; ++p) {
    const bool last = p == passes - 1;
    uint64_t* kout  = last ? keys_sorted_out : s->keys[p & 1];  // the last pass writes the keys only if somebody wants them
    uint32_t* vout  = last ? perm_out : s->vals[p & 1];
    onesweep_kernel<RadixDigit><<<ntiles, OS_THREADS, OS_SMEM, e->stream>>>(kin, vin, kout, vout, n, RadixDigit{8 * p}, s->ghist + p * OS_RADIX,
                                                               s->status + size_t(p) * ntiles * OS_RADIX, tickets + p);
    e->launches++;
    kin = kout;
    vin = vout;
  }
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

// Multi-GPU sharded sort (the alternative to every rank sorting all n keys). Every rank holds the same
// keys. (1) 65536 evenly spaced keys are sorted and world-1 splitters read off at the quantiles — identical on every rank;
// (2) ONE sweep partitions (key, index) by owner rank = number of splitters <= key, stably; (3) this rank radix-sorts
// only its own key range (about n / world pairs); (4) the permutation segments are exchanged with one grouped NCCL
// broadcast per owner. Equal keys share an owner and every stage is stable, so perm_out is bit-identical to sort_pairs'.
// One small host synchronisation (the world segment sizes) per call.
int sort_pairs_sharded(nbx_engine* e, const uint64_t* keys_in, uint32_t n, int key_bits, uint32_t* perm_out) {
  NBX_TRY(sorter_create(e, n));
  Sorter* s        = static_cast<Sorter*>(e->sorter);
  const int world  = e->cfg.world_size, rank = e->cfg.rank;
  if (world < 2 || world > OS_MAXSPLIT + 1 || !s->samples || n < 4 * OS_SAMPLES) return sort_pairs(e, keys_in, n, key_bits, perm_out, nullptr);
  // (1) splitters
  sample_keys_kernel<<<OS_SAMPLES / 256, 256, 0, e->stream>>>(keys_in, n, OS_SAMPLES, s->samples);
  e->launches++;
  NBX_TRY(sort_pairs(e, s->samples, OS_SAMPLES, key_bits, s->sample_perm, s->samples + OS_SAMPLES));
  OwnerDigit od;
  {
    uint64_t h[OS_MAXSPLIT];
    for (int q = 0; q < OS_MAXSPLIT; ++q) h[q] = ~0ull;
    // the quantile keys are needed on the host to be passed by value: one tiny strided copy, one sync
    NBX_CUDA(cudaMemcpy2DAsync(h, sizeof(uint64_t), s->samples + OS_SAMPLES + OS_SAMPLES / world, sizeof(uint64_t) * (OS_SAMPLES / world),
                               sizeof(uint64_t), size_t(world - 1), cudaMemcpyDeviceToHost, e->stream));
    NBX_CUDA(cudaStreamSynchronize(e->stream));
    e->d2h += sizeof(uint64_t) * (world - 1);
    for (int q = 0; q < OS_MAXSPLIT; ++q) od.split[q] = h[q];
    od.nsplit = world - 1;
  }
  // (2) stable partition by owner into the sorter's second buffers
  const uint32_t ntiles = (n + OS_TILE - 1) / OS_TILE;
  uint32_t* tickets     = s->ghist + OS_MAXPASS * OS_RADIX;
  NBX_CUDA(cudaMemsetAsync(s->ghist, 0, sizeof(uint32_t) * (OS_MAXPASS * OS_RADIX + OS_MAXPASS), e->stream));
  const unsigned hgrid = std::min<unsigned>(ntiles, unsigned(e->sm_count) * 4);
  owner_histogram_kernel<<<hgrid, OS_THREADS, 0, e->stream>>>(keys_in, n, od, s->ghist, s->status, size_t(ntiles) * OS_RADIX);
  uint32_t counts[OS_MAXSPLIT + 1];
  NBX_CUDA(cudaMemcpyAsync(counts, s->ghist, sizeof(uint32_t) * world, cudaMemcpyDeviceToHost, e->stream));
  digit_scan_kernel<<<1, OS_RADIX, 0, e->stream>>>(s->ghist);
  onesweep_kernel<OwnerDigit><<<ntiles, OS_THREADS, OS_SMEM, e->stream>>>(keys_in, nullptr, s->keys[1], s->vals[1], n, od, s->ghist, s->status, tickets);
  e->launches += 3;
  NBX_CUDA(cudaStreamSynchronize(e->stream));
  e->d2h += sizeof(uint32_t) * world;
  size_t offs[OS_MAXSPLIT + 2];
  offs[0] = 0;
  for (int q = 0; q < world; ++q) offs[q + 1] = offs[q] + counts[q];
  if (offs[world] != n) return fail(NBX_ERR_STATE, "sharded sort: partition sizes do not add up");
  // (3) my key range: sort_pairs ping-pongs through keys[0] / keys[1] from the front; its first pass has consumed the
  // segment before anything is written over the partition buffer
  const uint32_t cnt = counts[rank];
  if (cnt) NBX_TRY(sort_pairs(e, s->keys[1] + offs[rank], cnt, key_bits, perm_out + offs[rank], nullptr, s->vals[1] + offs[rank]));
  // (4) everybody gets every segment
  size_t boff[OS_MAXSPLIT + 1], bcnt[OS_MAXSPLIT + 1];
  for (int q = 0; q < world; ++q) { boff[q] = offs[q] * sizeof(uint32_t); bcnt[q] = size_t(counts[q]) * sizeof(uint32_t); }
  return comm_allgatherv(e, perm_out, boff, bcnt);
}

}  // namespace nbx
