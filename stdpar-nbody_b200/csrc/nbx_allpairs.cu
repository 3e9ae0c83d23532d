// nbx_allpairs.cu — O(n^2) force kernels + leapfrog for sm_100a.
//
//   all_pairs_kernel      : reference all_pairs_force            (src/all_pairs.h:14-27)   [K1]
//   collapsed_kernel      : reference all_pairs_collapsed_force  (src/all_pairs.h:29-50)   [K2]
//   accelerate_kernel     : reference System::accelerate_step    (src/system.h:52-60)      [K3]
//   energy kernels        : reference System::calc_energies      (src/system.h:62-79)      [K4]
//
// Design: targets are register-blocked (TI per thread), the j bodies stream through
// shared memory in TILE-sized chunks staged by 1-D TMA bulk copies (cp.async.bulk + mbarrier, STAGES deep) and
// are read back as broadcast LDS.128. The pair interaction is FMA/SFU work only: the kernel is bound by the
// MUFU pipe (sqrt + rcp per pair), tensor cores are deliberately not used. The j range is split over gridDim.y
// so that the grid has >= ~32 waves of CTAs; partial sums are combined in a FIXED order by the last-arriving CTA
// of each target block, which also applies the leapfrog update (fused epilogue, double-buffered positions).
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include "nbx_internal.cuh"
#include "nbx_math.cuh"
#include "nbx_device.cuh"

namespace nbx {

template <typename T, int D>
__global__ void accelerate_kernel(LeapArgs<T> p, uint32_t tb, uint32_t te) {
  uint32_t i = tb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= te) return;
  leapfrog_body<T, D>(p, i, p.a[i]);
}

// ---- all-pairs ------------------------------------------------------------------------------------------------
template <typename T>
struct AllPairsArgs {
  const vec4_t<T>* xm;   // sources (n_pad + TILE zero-mass padding records)
  uint32_t n;            // number of sources
  uint32_t tb, te;       // targets [tb, te)
  uint32_t chunk;        // stride of one j-split slab in `partial`
  uint32_t tiles_total;  // ceil(n / TILE)
  uint32_t tiles_per_split;
  vec4_t<T>* partial;    // [gridDim.y][chunk]
  uint32_t* tickets;     // [gridDim.x]
  T c;                   // System::constant
  int fuse;              // 1: apply the leapfrog in the epilogue
  LeapArgs<T> leap;      // leap.a is also the destination of the acceleration
};

// PACKED (float, TI even): the pair arithmetic of TWO targets per FP32x2 instruction (FFMA2/FADD2/FMUL2), the j body
// duplicated into both halves. Per target the operations and their order are those of the scalar loop, so the result is
// bit-identical; only the issue slots of the FMA-pipe work halve.
template <typename T, int D, int TI, int BLOCK, int TILE, int STAGES, int MINB, bool PACKED = false>
__global__ void __launch_bounds__(BLOCK, MINB) all_pairs_kernel(AllPairsArgs<T> p) {
  static_assert(!PACKED || (sizeof(T) == 4 && TI % 2 == 0), "packed arithmetic is FP32x2 over pairs of targets");
  using V4 = vec4_t<T>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  V4* tiles      = reinterpret_cast<V4*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + size_t(STAGES) * TILE * sizeof(V4));
  __shared__ uint32_t s_ticket;

  const int tid         = threadIdx.x;
  const uint32_t i_base = p.tb + blockIdx.x * (BLOCK * TI) + tid;

  T xi[TI], yi[TI], zi[TI], ax[TI], ay[TI], az[TI];
#pragma unroll
  for (int t = 0; t < TI; ++t) {
    uint32_t i = i_base + t * BLOCK;
    V4 b       = p.xm[i < p.te ? i : p.tb];
    xi[t] = b.x; yi[t] = b.y; zi[t] = b.z;
    ax[t] = ay[t] = az[t] = T(0);
  }

  const uint32_t tile_begin = blockIdx.y * p.tiles_per_split;
  uint32_t tile_end         = tile_begin + p.tiles_per_split;
  if (tile_end > p.tiles_total) tile_end = p.tiles_total;
  const int ntiles = tile_end > tile_begin ? int(tile_end - tile_begin) : 0;

  constexpr uint32_t TILE_BYTES = TILE * sizeof(V4);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < STAGES && s < ntiles; ++s) {
      mbar_expect_tx(&bars[s], TILE_BYTES);
      tma_load_1d(tiles + size_t(s) * TILE, p.xm + size_t(tile_begin + s) * TILE, TILE_BYTES, &bars[s]);
    }
  }

  for (int k = 0; k < ntiles; ++k) {
    const int stage = k % STAGES;
    mbar_wait(&bars[stage], (k / STAGES) & 1);
    const V4* tile = tiles + size_t(stage) * TILE;
    if constexpr (PACKED) {
      constexpr int TP = TI / 2;
      float2 nx2[TP], ny2[TP], nz2[TP], ax2[TP], ay2[TP], az2[TP];
#pragma unroll
      for (int q = 0; q < TP; ++q) {
        nx2[q] = make_float2(-xi[2 * q], -xi[2 * q + 1]); ny2[q] = make_float2(-yi[2 * q], -yi[2 * q + 1]);
        nz2[q] = make_float2(-zi[2 * q], -zi[2 * q + 1]);
        ax2[q] = make_float2(ax[2 * q], ax[2 * q + 1]); ay2[q] = make_float2(ay[2 * q], ay[2 * q + 1]);
        az2[q] = make_float2(az[2 * q], az[2 * q + 1]);
      }
#pragma unroll 4
      for (int j = 0; j < TILE; ++j) {
        const V4 b = tile[j];
        const float2 bx = make_float2(b.x, b.x), by = make_float2(b.y, b.y), bz = make_float2(b.z, b.z), bm = make_float2(b.w, b.w);
#pragma unroll
        for (int q = 0; q < TP; ++q) {
          const float2 dx = __fadd2_rn(bx, nx2[q]), dy = __fadd2_rn(by, ny2[q]);
          float2 d2 = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
          float2 dz = make_float2(0.f, 0.f);
          if (D == 3) { dz = __fadd2_rn(bz, nz2[q]); d2 = __ffma2_rn(dz, dz, d2); }
          float2 sq, inv;
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq.x) : "f"(d2.x));
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq.y) : "f"(d2.y));
          const float2 den = __ffma2_rn(d2, sq, make_float2(FLT_EPSILON, FLT_EPSILON));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(den.x));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(den.y));
          const float2 sm = __fmul2_rn(bm, inv);
          ax2[q] = __ffma2_rn(dx, sm, ax2[q]);
          ay2[q] = __ffma2_rn(dy, sm, ay2[q]);
          if (D == 3) az2[q] = __ffma2_rn(dz, sm, az2[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < TP; ++q) {
        ax[2 * q] = ax2[q].x; ax[2 * q + 1] = ax2[q].y; ay[2 * q] = ay2[q].x; ay[2 * q + 1] = ay2[q].y;
        az[2 * q] = az2[q].x; az[2 * q + 1] = az2[q].y;
      }
    } else
#pragma unroll 4
    for (int j = 0; j < TILE; ++j) {
      const V4 b = tile[j];  // broadcast LDS.128 (2x for double)
#pragma unroll
      for (int t = 0; t < TI; ++t) {
        T dx = b.x - xi[t];
        T dy = b.y - yi[t];
        T d2 = fma(dy, dy, sq_plus_tiny(dx));
        T dz = T(0);
        if (D == 3) {
          dz = b.z - zi[t];
          d2 = fma(dz, dz, d2);
        }
        T s   = b.w * inv_dist3_pos(d2);
        ax[t] = fma(dx, s, ax[t]);
        ay[t] = fma(dy, s, ay[t]);
        if (D == 3) az[t] = fma(dz, s, az[t]);
      }
    }
    __syncthreads();  // every warp is done with this stage: safe to overwrite it
    if (tid == 0 && k + STAGES < ntiles) {
      mbar_expect_tx(&bars[stage], TILE_BYTES);
      tma_load_1d(tiles + size_t(stage) * TILE, p.xm + size_t(tile_begin + k + STAGES) * TILE, TILE_BYTES,
                  &bars[stage]);
    }
  }

  // ---- epilogue: combine the j-splits in fixed order, scale by c, (optionally) leapfrog ------------------------
  const uint32_t nsplit = gridDim.y;
  if (nsplit > 1) {
#pragma unroll
    for (int t = 0; t < TI; ++t) {
      uint32_t i = i_base + t * BLOCK;
      if (i < p.te) stcg_v4(&p.partial[size_t(blockIdx.y) * p.chunk + (i - p.tb)], make_v4<T>(ax[t], ay[t], az[t], T(0)));
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&p.tickets[blockIdx.x], 1u);
    __syncthreads();
    if (s_ticket != nsplit - 1) return;
    __threadfence();
    if (tid == 0) p.tickets[blockIdx.x] = 0;  // self-reset for the next launch
#pragma unroll
    for (int t = 0; t < TI; ++t) {
      uint32_t i = i_base + t * BLOCK;
      ax[t] = ay[t] = az[t] = T(0);
      if (i < p.te) {
        for (uint32_t s = 0; s < nsplit; ++s) {
          V4 q = ldcg_v4(&p.partial[size_t(s) * p.chunk + (i - p.tb)]);
          ax[t] += q.x; ay[t] += q.y; az[t] += q.z;
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < TI; ++t) {
    uint32_t i = i_base + t * BLOCK;
    if (i >= p.te) continue;
    V4 anew = make_v4<T>(mul_rn(ax[t], p.c), mul_rn(ay[t], p.c), D == 3 ? mul_rn(az[t], p.c) : T(0), T(0));
    p.leap.a[i] = anew;
    if (p.fuse) leapfrog_body<T, D>(p.leap, i, anew);
  }
}

// ---- all-pairs-collapsed (pair-parallel) ----------------------------------------------------------------------
// Work-item p = j*n + i of the reference becomes: lanes <-> j, 8 rows i per warp, 8 warps (64 rows) per CTA; the j
// tile is staged in shared memory once per CTA. Each lane accumulates its partial row sums over the CTA's j range,
// a warp-shuffle butterfly reduces them and ONE relaxed atomic per row/component/CTA lands in a[i] — the same
// relaxed fetch_add the reference issues per pair (all_pairs.h:47-48). NC = 2 reproduces the reference's
// two-component accumulate; NC = 3 is the opt-in fix.
template <typename T>
struct CollapsedArgs {
  const vec4_t<T>* xm;
  uint32_t n, tb, te;
  uint32_t tiles_total, tiles_per_split;
  vec4_t<T>* a;
  T c;
};

template <typename T>
__global__ void collapsed_reset_kernel(vec4_t<T>* a, const vec4_t<T>* ao, uint32_t tb, uint32_t te, int nc) {
  uint32_t i = tb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= te) return;
  vec4_t<T> q = a[i], o = ao[i];
  // all_pairs.h:35-39: "reset to zero by subtracting old acceleration"
  q.x = add_rn(q.x, -o.x);
  q.y = add_rn(q.y, -o.y);
  if (nc == 3) q.z = add_rn(q.z, -o.z);
  a[i] = q;
}

template <typename T, int D, int NC, int TILE>
__global__ void __launch_bounds__(256) collapsed_kernel(CollapsedArgs<T> p) {
  using V4 = vec4_t<T>;
  constexpr int ROWS = 8;
  __shared__ V4 tile[TILE];
  __shared__ V4 rows[8 * ROWS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t row0 = p.tb + blockIdx.x * (8 * ROWS);
  if (tid < 8 * ROWS) {
    uint32_t i = row0 + tid;
    rows[tid]  = p.xm[i < p.te ? i : p.tb];
  }
  T acc[ROWS][NC];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int k = 0; k < NC; ++k) acc[r][k] = T(0);

  const uint32_t tile_begin = blockIdx.y * p.tiles_per_split;
  uint32_t tile_end         = tile_begin + p.tiles_per_split;
  if (tile_end > p.tiles_total) tile_end = p.tiles_total;
  for (uint32_t tk = tile_begin; tk < tile_end; ++tk) {
    __syncthreads();
    for (int q = tid; q < TILE; q += 256) tile[q] = p.xm[size_t(tk) * TILE + q];  // zero-mass padding past n
    __syncthreads();
#pragma unroll 2
    for (int jj = lane; jj < TILE; jj += 32) {
      const V4 b = tile[jj];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const V4 ri = rows[warp * ROWS + r];  // broadcast
        T dx = b.x - ri.x;
        T dy = b.y - ri.y;
        T d2 = fma(dy, dy, sq_plus_tiny(dx));
        T dz = T(0);
        if (D == 3) {
          dz = b.z - ri.z;
          d2 = fma(dz, dz, d2);
        }
        T s       = b.w * inv_dist3_pos(d2);
        acc[r][0] = fma(dx, s, acc[r][0]);
        acc[r][1] = fma(dy, s, acc[r][1]);
        if (NC == 3) acc[r][2] = fma(dz, s, acc[r][2]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      T vsum = acc[r][k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, off);
      acc[r][k] = vsum;
    }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      uint32_t i = row0 + warp * ROWS + r;
      if (i < p.te) {
        T* dst = reinterpret_cast<T*>(&p.a[i]);
#pragma unroll
        for (int k = 0; k < NC; ++k) atomicAdd(dst + k, p.c * acc[r][k]);  // RED.ADD relaxed
      }
    }
  }
}

// ---- energies (system.h:62-79) --------------------------------------------------------------------------------
// kinetic = 1/2 sum m v^2 ; gravitational = -1/2 G sum_i sum_{j != i} m_i m_j / (|x_i - x_j| + eps)   (dist, vec.h:243-246).
// The pair term is symmetric, so only the strictly upper triangle is evaluated: the CTA of target tile b sweeps the source
// tiles jt >= b (in its own tile only the pairs j > i) and the sum is doubled. Per pair: d2, 1/(sqrt(d2)+eps) from MUFU
// seeds (nbx_math.cuh), one FMA into a per-tile partial in T; the partials are accumulated in double. CTAs are launched heaviest first.
template <typename T, int D>
__global__ void __launch_bounds__(256) energy_kernel(const vec4_t<T>* __restrict__ xm, const vec4_t<T>* __restrict__ v, uint32_t n,
                                                      double* out /* [0]=sum m v^2, [1]=sum_i sum_j!=i mi mj / dist */) {
  using V4 = vec4_t<T>;
  __shared__ V4 tile[256];
  __shared__ double red[2][8];
  const uint32_t b = blockIdx.x, tid = threadIdx.x;
  const uint32_t i = b * 256 + tid;
  const V4 bi      = xm[i < n ? i : n - 1];
  const V4 vi      = v[i < n ? i : n - 1];
  double ke    = i < n ? double(bi.w) * (double(vi.x) * vi.x + double(vi.y) * vi.y + double(vi.z) * vi.z) : 0.0;
  double total = 0;
  for (uint32_t j0 = b * 256; j0 < n; j0 += 256) {
    __syncthreads();
    const uint32_t j = j0 + tid;
    tile[tid]        = j < n ? xm[j] : make_v4<T>(0, 0, 0, 0);  // zero mass: contributes exactly 0
    __syncthreads();
    // four interleaved partial sums in T per tile, then double: a heavy source (a galaxy centre) swamps the small terms
    // added after it in its own partial only
    T part[4] = {0, 0, 0, 0};
    if (j0 == b * 256) {  // own tile: strictly upper triangle
#pragma unroll 4
      for (uint32_t q = 0; q < 256; ++q) {
        const V4 s = tile[q];
        const T dx = bi.x - s.x, dy = bi.y - s.y, dz = D == 3 ? bi.z - s.z : T(0);
        T d2 = fma(dy, dy, sq_plus_tiny(dx));
        if (D == 3) d2 = fma(dz, dz, d2);
        const T t = s.w * inv_dist_eps(d2);
        part[q & 3] += q > tid ? t : T(0);
      }
    } else {
#pragma unroll 8
      for (uint32_t q = 0; q < 256; ++q) {
        const V4 s = tile[q];
        const T dx = bi.x - s.x, dy = bi.y - s.y, dz = D == 3 ? bi.z - s.z : T(0);
        T d2 = fma(dy, dy, sq_plus_tiny(dx));
        if (D == 3) d2 = fma(dz, dz, d2);
        part[q & 3] = fma(s.w, inv_dist_eps(d2), part[q & 3]);
      }
    }
    total += (double(part[0]) + double(part[1])) + (double(part[2]) + double(part[3]));
  }
  total = i < n ? 2.0 * double(bi.w) * total : 0.0;
  for (int off = 16; off > 0; off >>= 1) {
    ke += __shfl_xor_sync(0xffffffffu, ke, off);
    total += __shfl_xor_sync(0xffffffffu, total, off);
  }
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = ke;
    red[1][tid >> 5] = total;
  }
  __syncthreads();
  if (tid == 0) {
    double k2 = 0, g2 = 0;
    for (int w = 0; w < 8; ++w) { k2 += red[0][w]; g2 += red[1][w]; }
    atomicAdd(&out[0], k2);
    atomicAdd(&out[1], g2);
  }
}

// ---- FMA-pipe peak microbenchmarks (roofline denominators; MEASURED_PEAKS.json has none for FP32/FP64) ---------
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
  T r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
      r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
  }
  T s = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
  if (s == T(123456789)) out[0] = s;
}

int measure_fma_peak(int device, int precision, double* tflops) {
  if (!tflops) return fail(NBX_ERR_INVALID, "tflops is NULL");
  NBX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  NBX_CUDA(cudaGetDeviceProperties(&prop, device));
  void* out = nullptr;
  NBX_CUDA(cudaMalloc(&out, 64));
  cudaEvent_t e0, e1;
  NBX_CUDA(cudaEventCreate(&e0));
  NBX_CUDA(cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  const int iters  = precision == NBX_F32 ? 20000 : 10000;
  double best      = 0;
  for (int rep = 0; rep < 4; ++rep) {
    NBX_CUDA(cudaEventRecord(e0));
    if (precision == NBX_F32) fma_peak_kernel<float><<<blocks, threads>>>((float*)out, iters, 1.0000001f, 1e-9f);
    else fma_peak_kernel<double><<<blocks, threads>>>((double*)out, iters, 1.0000001, 1e-9);
    NBX_CUDA(cudaEventRecord(e1));
    NBX_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    NBX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * double(blocks) * threads * double(iters) * 64.0;
    double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return NBX_OK;
}

// ---- host side ------------------------------------------------------------------------------------------------
template <typename T>
static LeapArgs<T> make_leap(nbx_engine* e, bool to_next) {
  LeapArgs<T> l;
  l.xm_in  = static_cast<const vec4_t<T>*>(e->xm[e->cur]);
  l.xm_out = static_cast<vec4_t<T>*>(to_next ? e->xm[e->cur ^ 1] : e->xm[e->cur]);
  l.v      = static_cast<vec4_t<T>*>(e->v);
  l.a      = static_cast<vec4_t<T>*>(e->a);
  l.ao     = static_cast<vec4_t<T>*>(e->ao);
  l.dt     = T(e->cfg.dt);
  return l;
}

constexpr int AP_STAGES = 4;  // TILE (bodies per shared-memory tile) is 512, or 128 for small n (more CTAs); the position
                              // buffers carry >= 1024 zero-mass padding records, so whole tiles can always be loaded

template <typename T, int D, int TI, int BLOCK, int MINB, int AP_TILE, bool PACKED = false>
static int launch_all_pairs_cfg(nbx_engine* e, bool fuse, uint32_t nsplit, uint32_t tiles_per_split) {
  auto kern = all_pairs_kernel<T, D, TI, BLOCK, AP_TILE, AP_STAGES, MINB, PACKED>;
  const size_t smem = size_t(AP_STAGES) * AP_TILE * sizeof(vec4_t<T>) + AP_STAGES * sizeof(uint64_t);
  NBX_TRY(ensure_dynamic_smem(e, kern, smem));
  const uint32_t nt      = e->te - e->tb;
  const uint32_t iblocks = (nt + BLOCK * TI - 1) / (BLOCK * TI);
  if (nsplit > 1) {
    size_t need = size_t(nsplit) * e->chunk * sizeof(vec4_t<T>);
    if (need > e->partial_bytes) {
      if (e->partial) cudaFree(e->partial);
      e->partial = nullptr;
      NBX_CUDA(cudaMalloc(&e->partial, need));
      e->partial_bytes = need;
    }
    if (iblocks > e->tickets_count) {
      if (e->tickets) cudaFree(e->tickets);
      e->tickets = nullptr;
      NBX_CUDA(cudaMalloc(&e->tickets, sizeof(uint32_t) * iblocks));
      NBX_CUDA(cudaMemsetAsync(e->tickets, 0, sizeof(uint32_t) * iblocks, e->stream));
      e->tickets_count = iblocks;
    }
  }
  AllPairsArgs<T> p;
  p.xm              = static_cast<const vec4_t<T>*>(e->xm[e->cur]);
  p.n               = e->n;
  p.tb              = e->tb;
  p.te              = e->te;
  p.chunk           = e->chunk;
  p.tiles_total     = (e->n + AP_TILE - 1) / AP_TILE;
  p.tiles_per_split = tiles_per_split;
  p.partial         = static_cast<vec4_t<T>*>(e->partial);
  p.tickets         = e->tickets;
  p.c               = T(e->cfg.G);
  p.fuse            = fuse ? 1 : 0;
  p.leap            = make_leap<T>(e, /*to_next=*/true);
  dim3 grid(iblocks, nsplit);
  kern<<<grid, BLOCK, smem, e->stream>>>(p);
  e->launches++;
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

template <typename T, int D, int AP_TILE>
static int launch_all_pairs(nbx_engine* e, bool fuse) {
  const uint32_t nt = e->te - e->tb;
  if (nt == 0) return NBX_OK;
  const uint32_t tiles_total = (e->n + AP_TILE - 1) / AP_TILE;
  // pick the target register blocking so that small problems still fill the chip
  constexpr int TI_MAX = sizeof(T) == 4 ? 4 : 2;
  int ti               = TI_MAX;
  const uint32_t want  = uint32_t(e->sm_count) * 4;
  while (ti > 1 && ((nt + 256 * ti - 1) / (256 * ti)) * tiles_total < want * 4) ti >>= 1;
  if (const char* f = getenv("NBX_AP_TI")) {  // experiments: force the target blocking
    const int v = atoi(f);
    if (v == 1 || v == 2 || (v == 4 && TI_MAX >= 4)) ti = v;
  }
  const int block        = 256;
  const uint32_t iblocks = (nt + block * ti - 1) / (block * ti);
  // j-split: aim for >= 32 waves of CTAs (3 resident per SM) while keeping >= 4 tiles per CTA
  uint32_t target_ctas = uint32_t(e->sm_count) * 3 * 32;
  uint32_t nsplit      = (target_ctas + iblocks - 1) / iblocks;
  uint32_t max_split   = tiles_total / 4 ? tiles_total / 4 : 1;
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  uint32_t tps = (tiles_total + nsplit - 1) / nsplit;
  nsplit       = (tiles_total + tps - 1) / tps;
  if constexpr (sizeof(T) == 4) {
    const char* pk = getenv("NBX_AP_PACKED");  // experiments: FP32x2 arithmetic over pairs of targets (bit-identical results)
    if (pk && atoi(pk)) {
      if (ti == 4) return launch_all_pairs_cfg<T, D, 4, 256, 3, AP_TILE, true>(e, fuse, nsplit, tps);
      if (ti == 2) return launch_all_pairs_cfg<T, D, 2, 256, 3, AP_TILE, true>(e, fuse, nsplit, tps);
    }
    if (ti == 4) return launch_all_pairs_cfg<T, D, 4, 256, 3, AP_TILE>(e, fuse, nsplit, tps);
    if (ti == 2) return launch_all_pairs_cfg<T, D, 2, 256, 3, AP_TILE>(e, fuse, nsplit, tps);
    return launch_all_pairs_cfg<T, D, 1, 256, 3, AP_TILE>(e, fuse, nsplit, tps);
  } else {
    if (ti == 2) return launch_all_pairs_cfg<T, D, 2, 256, 2, AP_TILE>(e, fuse, nsplit, tps);
    return launch_all_pairs_cfg<T, D, 1, 256, 2, AP_TILE>(e, fuse, nsplit, tps);
  }
}

int all_pairs_force(nbx_engine* e, bool fuse_integrate) {
  PhaseTimer pt(e, PH_FORCE);
  int rc;
  if (all_pairs_sym_enabled(e)) {
    rc = all_pairs_sym_force(e, fuse_integrate);
    if (rc == NBX_OK && fuse_integrate) e->cur ^= 1;
    // the symmetric kernel's partial-sum buffer does not fit: on one GPU fall back to the ordered sweep for good (same
    // result up to summation order); on several GPUs every rank must take the same path, so the error stands
    if (rc != NBX_ERR_CAPACITY || e->cfg.world_size > 1) return rc;
    e->sym_unavailable = true;
  }
  const bool small = e->n < 32768;  // 128-body tiles give small problems 4x more CTAs to spread over the 148 SMs
  if (e->prec == 4) {
    if (e->dim == 2) rc = small ? launch_all_pairs<float, 2, 128>(e, fuse_integrate) : launch_all_pairs<float, 2, 512>(e, fuse_integrate);
    else rc = small ? launch_all_pairs<float, 3, 128>(e, fuse_integrate) : launch_all_pairs<float, 3, 512>(e, fuse_integrate);
  } else {
    if (e->dim == 2) rc = small ? launch_all_pairs<double, 2, 128>(e, fuse_integrate) : launch_all_pairs<double, 2, 512>(e, fuse_integrate);
    else rc = small ? launch_all_pairs<double, 3, 128>(e, fuse_integrate) : launch_all_pairs<double, 3, 512>(e, fuse_integrate);
  }
  if (rc == NBX_OK && fuse_integrate) e->cur ^= 1;
  // multi-GPU, ordered kernel: every rank computed a[tb, te) — gather the shards so that `a` is complete everywhere
  if (rc == NBX_OK && e->cfg.world_size > 1) {
    if (fuse_integrate) return fail(NBX_ERR_STATE, "ordered all-pairs on several GPUs integrates after the all-gather of a");
    rc = comm_allgather(e, e->a);
  }
  return rc;
}

template <typename T, int D>
static int launch_collapsed(nbx_engine* e) {
  const uint32_t nt = e->te - e->tb;
  if (nt == 0) return NBX_OK;
  constexpr int TILE = 512;
  const int nc       = (e->cfg.flags & NBX_FLAG_COLLAPSED_FIX_Z) && D == 3 ? 3 : 2;
  collapsed_reset_kernel<T><<<(nt + 255) / 256, 256, 0, e->stream>>>(static_cast<vec4_t<T>*>(e->a),
                                                                     static_cast<const vec4_t<T>*>(e->ao), e->tb, e->te, nc);
  e->launches++;
  CollapsedArgs<T> p;
  p.xm          = static_cast<const vec4_t<T>*>(e->xm[e->cur]);
  p.n           = e->n;
  p.tb          = e->tb;
  p.te          = e->te;
  p.tiles_total = (e->n + TILE - 1) / TILE;
  p.a           = static_cast<vec4_t<T>*>(e->a);
  p.c           = T(e->cfg.G);
  const uint32_t rblocks = (nt + 63) / 64;
  uint32_t target_ctas   = uint32_t(e->sm_count) * 4 * 16;
  uint32_t nsplit        = (target_ctas + rblocks - 1) / rblocks;
  uint32_t max_split     = p.tiles_total / 2 ? p.tiles_total / 2 : 1;
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  p.tiles_per_split = (p.tiles_total + nsplit - 1) / nsplit;
  nsplit            = (p.tiles_total + p.tiles_per_split - 1) / p.tiles_per_split;
  dim3 grid(rblocks, nsplit);
  if (nc == 3) collapsed_kernel<T, D, 3, TILE><<<grid, 256, 0, e->stream>>>(p);
  else collapsed_kernel<T, D, 2, TILE><<<grid, 256, 0, e->stream>>>(p);
  e->launches++;
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

int all_pairs_collapsed_force(nbx_engine* e) {
  PhaseTimer pt(e, PH_FORCE);
  if (all_pairs_sym_enabled(e)) {
    // large n: the block-pair units of the symmetric kernel ARE the pair-parallel decomposition, with the symmetry the
    // reference's TODO asks for (all_pairs.h:41-42); the collapsed semantics (2 components, a -= ao reset) live in the finish
    const int nc = (e->cfg.flags & NBX_FLAG_COLLAPSED_FIX_Z) && e->dim == 3 ? 3 : 2;
    const int rcs = all_pairs_sym_force(e, false, nc);
    if (rcs != NBX_ERR_CAPACITY || e->cfg.world_size > 1) return rcs;
    e->sym_unavailable = true;  // buffer does not fit: the pair-parallel kernel below serves instead
  }
  int rc;
  if (e->prec == 4) rc = e->dim == 2 ? launch_collapsed<float, 2>(e) : launch_collapsed<float, 3>(e);
  else rc = e->dim == 2 ? launch_collapsed<double, 2>(e) : launch_collapsed<double, 3>(e);
  // multi-GPU: every rank updated a[tb, te) (reset + relaxed adds) — gather the shards, `a` is complete everywhere
  if (rc == NBX_OK && e->cfg.world_size > 1) rc = comm_allgather(e, e->a);
  return rc;
}

template <typename T, int D>
static int launch_accelerate(nbx_engine* e, uint32_t tb, uint32_t te) {
  const uint32_t nt = te - tb;
  if (nt == 0) return NBX_OK;
  accelerate_kernel<T, D><<<(nt + 255) / 256, 256, 0, e->stream>>>(make_leap<T>(e, /*to_next=*/false), tb, te);
  e->launches++;
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

int accelerate_range(nbx_engine* e, uint32_t tb, uint32_t te) {
  PhaseTimer pt(e, PH_ACCEL);
  if (e->prec == 4) return e->dim == 2 ? launch_accelerate<float, 2>(e, tb, te) : launch_accelerate<float, 3>(e, tb, te);
  return e->dim == 2 ? launch_accelerate<double, 2>(e, tb, te) : launch_accelerate<double, 3>(e, tb, te);
}

// The whole state_t is replicated on every rank for every algorithm: the force functions leave the complete
// acceleration array everywhere (all-gather of the target shards, or all-reduce of the symmetric kernel's per-rank sums),
// so every rank integrates all bodies and v, a, ao never go stale outside a rank's shard.
int accelerate_step(nbx_engine* e) { return accelerate_range(e, 0, e->n); }

template <typename T, int D>
static int launch_energies(nbx_engine* e, double* out_dev) {
  energy_kernel<T, D><<<(e->n + 255) / 256, 256, 0, e->stream>>>(static_cast<const vec4_t<T>*>(e->xm[e->cur]),
                                                                 static_cast<const vec4_t<T>*>(e->v), e->n, out_dev);
  e->launches++;
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

int calc_energies(nbx_engine* e, double* kinetic, double* grav) {
  if (!e->energy_out) NBX_CUDA(cudaMalloc(&e->energy_out, 2 * sizeof(double)));
  double* out = e->energy_out;
  NBX_CUDA(cudaMemsetAsync(out, 0, 2 * sizeof(double), e->stream));
  int rc;
  if (e->prec == 4) rc = e->dim == 2 ? launch_energies<float, 2>(e, out) : launch_energies<float, 3>(e, out);
  else rc = e->dim == 2 ? launch_energies<double, 2>(e, out) : launch_energies<double, 3>(e, out);
  if (rc != NBX_OK) return rc;
  double h[2] = {0, 0};
  NBX_CUDA(cudaMemcpyAsync(h, out, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
  NBX_CUDA(cudaStreamSynchronize(e->stream));
  e->d2h += sizeof(h);
  if (kinetic) *kinetic = 0.5 * h[0];               // system.h:64-66
  if (grav) *grav = -0.5 * e->cfg.G * h[1];         // system.h:67-77
  return NBX_OK;
}

}  // namespace nbx
