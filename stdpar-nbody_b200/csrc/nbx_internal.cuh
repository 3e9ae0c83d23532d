// nbx_internal.cuh — engine state shared by the translation units of libnbx.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <set>
#include <string>

#include "nbx.h"

namespace nbx {

// ---- error plumbing ------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define NBX_CUDA(call)                                                                             \
  do {                                                                                             \
    cudaError_t err__ = (call);                                                                    \
    if (err__ != cudaSuccess)                                                                      \
      return ::nbx::fail(NBX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));     \
  } while (0)

#define NBX_TRY(call)            \
  do {                           \
    int rc__ = (call);           \
    if (rc__ != NBX_OK) return rc__; \
  } while (0)

// ---- packed 4-vectors: the HBM record of every per-body quantity ----------------------------------------------
// pos/mass record: (x, y, z, m) — z = 0 in 2-D; velocity/acceleration records: (vx, vy, vz, 0).
// float -> 16 B (one LDG/LDS.128), double -> 32 B (two).
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  using type = float4;
};
template <>
struct Vec4<double> {
  using type = double4;
};
template <typename T>
using vec4_t = typename Vec4<T>::type;

template <typename T>
__host__ __device__ inline vec4_t<T> make_v4(T x, T y, T z, T w) {
  vec4_t<T> r;
  r.x = x; r.y = y; r.z = z; r.w = w;
  return r;
}

enum PhaseSlot { PH_FORCE = 0, PH_ACCEL, PH_BBOX, PH_SORT, PH_BUILD, PH_MONO, PH_TRAVERSE, PH_COMM, PH_COUNT };

}  // namespace nbx

// The opaque handle of nbx.h
struct nbx_engine {
  nbx_config cfg{};
  int dim = 3, prec = 4, algo = 0;
  uint32_t n = 0;        // bodies
  uint32_t chunk = 0;    // targets per rank = ceil(n / world)
  uint32_t n_pad = 0;    // chunk * world (records allocated for the position buffers; padding has m = 0)
  uint32_t tb = 0, te = 0;  // this rank's targets [tb, te)
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;

  // body state, vec4 records (4*prec bytes each)
  void* xm[2] = {nullptr, nullptr};  // double-buffered (x,y,z,m): force reads xm[cur], fused leapfrog writes xm[cur^1]
  int cur = 0;
  void* v = nullptr;
  void* a = nullptr;
  void* ao = nullptr;
  // alternates for the BVH permutation (swap after the gather)
  void* v_alt = nullptr;
  void* a_alt = nullptr;
  void* ao_alt = nullptr;

  // all-pairs scratch
  void* partial = nullptr;     // [jsplit][chunk] vec4 partial accelerations
  size_t partial_bytes = 0;
  uint32_t* tickets = nullptr; // per i-block arrival counters (self-resetting)
  size_t tickets_count = 0;
  double* energy_out = nullptr;  // [2] result slot of calc_energies (allocated on first use)

  // host<->device staging for upload/download (AoS <-> vec4 conversion happens on the device)
  void* stage = nullptr;
  size_t stage_bytes = 0;
  // Saver streaming: two snapshots in flight (device staging + pinned host buffer + "copy done" event each)
  cudaStream_t copy_stream = nullptr;
  void* snap_dev[2]        = {nullptr, nullptr};
  void* snap_host[2]       = {nullptr, nullptr};
  cudaEvent_t snap_ready[2] = {nullptr, nullptr};  // unpack kernel finished on the main stream
  cudaEvent_t snap_done[2]  = {nullptr, nullptr};  // D2H finished on the copy stream
  int snap_head = 0, snap_count = 0;               // ring of pending snapshots

  // tree state lives in opaque per-algorithm blocks owned by nbx_bvh.cu / nbx_octree.cu
  void* bvh = nullptr;
  void* octree = nullptr;
  void* sorter = nullptr;
  void* sym = nullptr;  // symmetric all-pairs state (nbx_allpairs_sym.cu)
  bool sym_unavailable = false;  // its partial-sum buffer did not fit (single GPU): the ordered kernel serves instead

  // NCCL (multi-GPU)
  void* comm = nullptr;
  // peer-memory exchange of the BVH accelerations: the two local buffers `a` alternates between and, once imported, the
  // other ranks' views of theirs (CUDA IPC), indexed [buffer][rank]
  void* own_a[2]       = {nullptr, nullptr};
  void* peer_a[2][16]  = {};
  bool peers_ready     = false;
  int* barrier_scratch = nullptr;

  // CUDA graph of one BVH time step (single GPU: the step has no host decision). The step flips the position buffer
  // and swaps v/a/ao with their alternates, so there is one graph per parity of `cur`.
  bool use_graph = true;  // NBX_GRAPH=0 (read when the engine is created) disables the replay
  cudaGraphExec_t step_graph[2] = {nullptr, nullptr};
  uint64_t step_graph_launches[2] = {0, 0};
  const void* step_graph_v[2] = {nullptr, nullptr};  // e->v when the graph was captured (must match at replay)

  // kernels whose >48 KB dynamic shared memory opt-in has been set for THIS engine's device (the attribute is
  // per device, and engines of several devices can live in one process)
  std::set<const void*> smem_opt_in;

  // counters / timing
  uint64_t launches = 0, h2d = 0, d2h = 0;
  bool phase_timing = false;
  cudaEvent_t ph_ev[nbx::PH_COUNT][2] = {};
  bool ph_used[nbx::PH_COUNT] = {};
  float ph_ms[nbx::PH_COUNT] = {};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace nbx {

struct PhaseTimer {
  nbx_engine* e;
  int slot;
  PhaseTimer(nbx_engine* e_, int slot_) : e(e_), slot(slot_) {
    if (e->phase_timing) cudaEventRecord(e->ph_ev[slot][0], e->stream);
  }
  ~PhaseTimer() {
    if (e->phase_timing) {
      cudaEventRecord(e->ph_ev[slot][1], e->stream);
      e->ph_used[slot] = true;
    }
  }
};

inline size_t rec_bytes(const nbx_engine* e) { return size_t(4) * e->prec; }

template <typename K>
inline int ensure_dynamic_smem(nbx_engine* e, K kernel, size_t bytes) {
  const void* key = reinterpret_cast<const void*>(kernel);
  if (e->smem_opt_in.count(key)) return NBX_OK;
  NBX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  e->smem_opt_in.insert(key);
  return NBX_OK;
}

// ---- implemented per translation unit -------------------------------------------------------------------------
// nbx_allpairs.cu
int all_pairs_force(nbx_engine* e, bool fuse_integrate);
int all_pairs_collapsed_force(nbx_engine* e);
int accelerate_step(nbx_engine* e);                               // over this rank's targets [tb, te)
int accelerate_range(nbx_engine* e, uint32_t begin, uint32_t end);  // over [begin, end)
int calc_energies(nbx_engine* e, double* kinetic, double* grav);
int measure_fma_peak(int device, int precision, double* tflops);
// nbx_allpairs_sym.cu : Newton's-third-law variant of all_pairs_force
bool all_pairs_sym_enabled(const nbx_engine* e);
uint32_t all_pairs_sym_block(uint32_t n, int world);
int all_pairs_sym_force(nbx_engine* e, bool fuse_integrate, int collapsed_nc = 0);
void all_pairs_sym_destroy(nbx_engine* e);
// nbx_sort.cu : stable LSD radix sort of (u64 key, u32 value) pairs
int sorter_create(nbx_engine* e, uint32_t n);
void sorter_destroy(nbx_engine* e);
// sorts keys_in (device, n) -> perm_out (device, n): perm_out[i] = index of the i-th smallest key (ties by index).
// keys_sorted_out (optional) receives the sorted keys. `key_bits` = number of significant low bits.
// `vals_in` (optional, must not alias the sorter's internal buffers or perm_out) replaces the implicit iota payload,
// so that a second, more significant key can be sorted on top of an earlier permutation (LSD over two words).
int sort_pairs(nbx_engine* e, const uint64_t* keys_in, uint32_t n, int key_bits, uint32_t* perm_out,
               uint64_t* keys_sorted_out, const uint32_t* vals_in = nullptr);
// multi-GPU: the ranks split the key range between them (sample splitters, one partition sweep, per-rank sort, exchange
// of the permutation segments); perm_out is bit-identical to sort_pairs'. Falls back to sort_pairs on one GPU / small n.
int sort_pairs_sharded(nbx_engine* e, const uint64_t* keys_in, uint32_t n, int key_bits, uint32_t* perm_out);
// nbx_bvh.cu
int bvh_create(nbx_engine* e);
void bvh_destroy(nbx_engine* e);
int bvh_bounding_box(nbx_engine* e);
int bvh_hilbert_sort(nbx_engine* e);
int bvh_build_tree(nbx_engine* e);
int bvh_compute_force(nbx_engine* e);
int bvh_get_bbox(nbx_engine* e, void* xmin, void* xmax);
void bvh_after_graph_replay(nbx_engine* e);                      // host-side effects of a replayed step (buffer swaps)
int octree_walk_width(const nbx_engine* e);                     // bodies per warp step of the octree walk (32, 16 or 8)
int bvh_walk_width(const nbx_engine* e);                        // bodies per warp step of the walk (lanes carrying bodies x bodies per lane)
int bvh_stats(nbx_engine* e, unsigned long long* dev_stats);     // counting variant of the walk: {visits, interactions, warp steps}
int bvh_get_keys(nbx_engine* e, uint64_t* keys, uint32_t* perm);
int bvh_get_nodes(nbx_engine* e, uint64_t* nnodes, void* node_m, void* bw, void* b);
// nbx_octree.cu
int octree_create(nbx_engine* e);
void octree_destroy(nbx_engine* e);
int octree_build(nbx_engine* e);
int octree_compute_force(nbx_engine* e);
int octree_stats(nbx_engine* e, unsigned long long* dev_stats);
int octree_check(nbx_engine* e);  // NBX_ERR_CAPACITY if the last build overflowed (syncs the stream)
int octree_get_root(nbx_engine* e, void* side, void* root_x, uint64_t* nodes_used);
int octree_get_canonical(nbx_engine* e, uint64_t* count, uint32_t* depth, uint64_t* path, uint32_t* kind,
                         void* monopole);
// nbx_comm.cu
int comm_unique_id(void* id128);
int comm_init_rank(nbx_engine* e, const void* id128);
int comm_allgather(nbx_engine* e, void* vec4_array);  // in place: rank r contributes records [r*chunk, (r+1)*chunk)
int comm_allreduce_sum(nbx_engine* e, void* buffer, size_t count);  // in place, `count` elements of the engine's precision
int comm_broadcast(nbx_engine* e, void* buffer, size_t bytes, int root);
int comm_barrier(nbx_engine* e);  // every rank's earlier work on its stream is complete (and its peer stores visible) when this returns on the stream
int peer_export(nbx_engine* e, void* handle128);
int peer_import(nbx_engine* e, const void* handles);
void peer_close(nbx_engine* e);
// in place: rank q owns bytes [offset[q], offset[q] + count[q]) of `buffer`; afterwards every rank holds all of them
int comm_allgatherv(nbx_engine* e, void* buffer, const size_t* offset, const size_t* count);
void comm_destroy(nbx_engine* e);

}  // namespace nbx
