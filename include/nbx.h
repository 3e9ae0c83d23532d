/* nbx.h — C ABI of the B200-native N-body force engine (libnbx.so).
 *
 * Drop-in boundary for the ONE data-parallel hot path of UoB-HPC/stdpar-nbody: the per-step
 * force + leapfrog pipeline behind `sim_func_t<T,N>` (reference src/main.cpp:16-17,58-64).
 * Plain pointers and sizes only; no C++/torch types. Every entry point names the reference
 * interface it replaces (paths relative to the reference root).
 *
 * State layout at the boundary is EXACTLY System<T,N>::state_t (src/system.h:41-50):
 *   m  : T[n]              x, v, a, ao : vec<T,N>[n]  (packed AoS, N*sizeof(T) stride)
 * T = float (precision NBX_F32) or double (NBX_F64); N = dim in {2,3}.
 * Internally the engine keeps (x,y,z,m)-packed 16/32-byte records in HBM.
 *
 * Threading: one caller thread per engine; every call returns after its work is enqueued on the engine's
 * CUDA stream, calls that copy to host memory return after the copy has completed. nbx_sync() blocks.
 * One exception: every octree build (nbx_octree_build, and each step of nbx_step on an NBX_OCTREE engine) waits once for
 * the device, after the tree records are emitted, to learn whether the bodies were separated within the key depth and
 * the cells fit — an unusable tree is reported as NBX_ERR_CAPACITY from that call, before any force kernel walks it.
 * Errors: every function returns NBX_OK (0) or a negative code; nbx_last_error() gives the message of the
 * last failure on the calling thread. There is NO CPU fallback: without a CUDA device nbx_create fails.
 */
#ifndef NBX_H
#define NBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBX_VERSION 100

enum nbx_status {
  NBX_OK              = 0,
  NBX_ERR_INVALID     = -1, /* bad argument / unsupported combination                                  */
  NBX_ERR_CUDA        = -2, /* a CUDA runtime call failed (message has the CUDA error string)           */
  NBX_ERR_NO_DEVICE   = -3, /* no usable CUDA device: the product has no CPU path                       */
  NBX_ERR_CAPACITY    = -4, /* octree deeper than the supported key depth / node capacity exceeded      */
  NBX_ERR_COMM        = -5, /* NCCL failure or communicator not initialised                             */
  NBX_ERR_STATE       = -6  /* call order violated (e.g. compute_force before build_tree)               */
};

/* src/arguments.h:16-21 SimulationAlgo (same order) */
enum nbx_algorithm { NBX_ALL_PAIRS = 0, NBX_ALL_PAIRS_COLLAPSED = 1, NBX_OCTREE = 2, NBX_BVH = 3 };
/* sizeof(T) */
enum nbx_precision { NBX_F32 = 4, NBX_F64 = 8 };

typedef struct nbx_engine nbx_engine;

typedef struct nbx_config {
  uint32_t struct_size;  /* = sizeof(nbx_config)                                                          */
  int32_t  dim;          /* -DDIM_SIZE (src/main.cpp:5-7): 2 or 3                                         */
  int32_t  precision;    /* --precision (src/arguments.h:58-69): NBX_F32 | NBX_F64                        */
  int32_t  algorithm;    /* --algorithm (src/arguments.h:70-86)                                           */
  uint32_t n;            /* System::size (src/system.h:14)                                                */
  int32_t  device;       /* CUDA device ordinal                                                           */
  double   dt;           /* System::dt   (src/system.h:16)                                                */
  double   G;            /* System::constant (src/system.h:17)                                            */
  double   theta;        /* --theta (src/arguments.h:31)                                                  */
  /* multi-GPU (world_size > 1, one engine per GPU): the state_t is REPLICATED on every rank; the force work of a step
   * is sharded (targets [rank*ceil(n/world), ...) for the ordered all-pairs / collapsed / tree walks, block-pair units
   * dealt round-robin for the symmetric all-pairs kernel), the accelerations are exchanged with one NCCL all-gather /
   * all-reduce per step and every rank integrates all bodies. world_size 1 = single GPU. */
  int32_t  rank;
  int32_t  world_size;
  uint32_t flags;        /* NBX_FLAG_* */
  uint32_t reserved;
} nbx_config;

#define NBX_FLAG_COLLAPSED_FIX_Z 0x1u /* non-default deviation: also accumulate z in all-pairs-collapsed */
#define NBX_FLAG_NO_FUSED_INTEGRATE 0x2u /* nbx_step runs force and leapfrog as separate kernels */
/* all-pairs evaluates each UNORDERED pair once (Newton's third law, the reference's own TODO at src/all_pairs.h:41-42)
 * when n >= 16384; results differ from the ordered sweep only by summation order. These flags override the choice: */
#define NBX_FLAG_ALLPAIRS_ORDERED 0x4u   /* always sweep all ordered pairs (the reference's loop structure) */
#define NBX_FLAG_ALLPAIRS_SYMMETRIC 0x8u /* use the symmetric kernel for any n */

const char* nbx_last_error(void);
int nbx_version(void);
/* number of CUDA devices visible (0 on a box without a GPU) */
int nbx_device_count(void);

/* replaces: System<T,N> construction + tree alloc (src/system.h:27-36, src/octree.h:42-51, src/bvh.h:147-164) */
int nbx_create(const nbx_config* cfg, nbx_engine** out);
/* replaces: ~System / bvh::dealloc (src/bvh.h:167-172) */
int nbx_destroy(nbx_engine* e);

/* replaces: System::state() pointer hand-off (src/system.h:47-50). Host arrays in state_t layout. Any pointer
 * may be NULL to skip that array. */
int nbx_upload(nbx_engine* e, const void* m, const void* x, const void* v, const void* a, const void* ao);
int nbx_download(nbx_engine* e, void* m, void* x, void* v, void* a, void* ao);
/* multi-GPU I/O: the pointers address the FULL host arrays as above, but only this rank's shard of bodies
 * [rank*ceil(n/world), (rank+1)*ceil(n/world)) is read / written. nbx_upload_shard then all-gathers the shards on the
 * device (NCCL over NVLink), so every rank still ends up with the complete replicated state while each sends only
 * 1/world of the bytes over PCIe; it is a collective: every rank of the communicator must call it with the same set of
 * non-NULL arrays. With world_size 1 both are nbx_upload / nbx_download. */
int nbx_upload_shard(nbx_engine* e, const void* m, const void* x, const void* v, const void* a, const void* ao);
int nbx_download_shard(nbx_engine* e, void* m, void* x, void* v, void* a, void* ao);

/* replaces: the `kernels()` lambda of run_all_pairs / run_octree / run_bvh (src/all_pairs.h:86-92,
 * src/octree.h:321-328, src/bvh.h:382-397): `steps` x (force + accelerate_step), device resident. */
int nbx_step(nbx_engine* e, uint32_t steps);
/* same, bracketed by CUDA events on the engine stream; *ms = device time of the `steps` steps */
int nbx_step_timed(nbx_engine* e, uint32_t steps, float* ms);
int nbx_sync(nbx_engine* e);

/* ---- per-phase entry points (the finer seams used inside the reference drivers) ---------------------------- */
/* all_pairs_force(System&)            src/all_pairs.h:14-27  */
int nbx_all_pairs_force(nbx_engine* e);
/* all_pairs_collapsed_force(System&)  src/all_pairs.h:29-50  */
int nbx_all_pairs_collapsed_force(nbx_engine* e);
/* System::accelerate_step()           src/system.h:52-60     */
int nbx_accelerate_step(nbx_engine* e);
/* System::calc_energies()             src/system.h:62-79  -> T kinetic, T gravitational (as double) */
int nbx_calc_energies(nbx_engine* e, double* kinetic, double* gravitational);

/* bounding_box(span)                  src/bvh.h:17-22   xmin/xmax: T[dim] host */
int nbx_bvh_bounding_box(nbx_engine* e, void* xmin, void* xmax);
/* hilbert_sort(System&, aabb)         src/bvh.h:25-96   (uses the box of the last nbx_bvh_bounding_box) */
int nbx_bvh_hilbert_sort(nbx_engine* e);
/* bvh::build_tree(System&)            src/bvh.h:175-244 */
int nbx_bvh_build_tree(nbx_engine* e);
/* bvh::compute_force(System&, theta)  src/bvh.h:251-324 */
int nbx_bvh_compute_force(nbx_engine* e);
/* artefacts of the last hilbert_sort: keys[n] BEFORE sorting (body order at entry), perm[n] (new[i]=old[perm[i]]) */
int nbx_bvh_get_keys(nbx_engine* e, uint64_t* keys, uint32_t* perm);
/* bvh::m / bw / b arrays (src/bvh.h:103-106): node_m T[nn*(dim+1)], bw T[nn], b T[nn*2*dim]; nn = *nnodes */
int nbx_bvh_get_nodes(nbx_engine* e, uint64_t* nnodes, void* node_m, void* bw, void* b);

/* octree::clear + compute_bounds + insert + compute_tree  src/octree.h:77-224 */
int nbx_octree_build(nbx_engine* e);
/* octree::compute_force(System&, theta)                   src/octree.h:258-263 */
int nbx_octree_compute_force(nbx_engine* e);
/* root cube of the last build (src/octree.h:93-112): side T, root_x T[dim]; nodes = 1 + 2^dim * internal cells
 * (what next_free_child_group would read, src/octree.h:152) */
int nbx_octree_get_root(nbx_engine* e, void* side, void* root_x, uint64_t* nodes_used);
/* numbering-independent canonical form of the last build, DFS child order, one record per NON-EMPTY node:
 * depth u32, path u64 (dim bits per level, root first), kind u32 (1 body leaf / 0 internal), monopole T[dim+1]
 * (x.., mass). Pass NULL arrays to query *count only. */
int nbx_octree_get_canonical(nbx_engine* e, uint64_t* count, uint32_t* depth, uint64_t* path, uint32_t* kind,
                             void* monopole);

/* ---- Saver streaming: positions.bin frames without stalling the GPU (replaces Saver::save_points, src/saving.h:110-114) ----
 * nbx_stream_positions_begin snapshots the CURRENT positions (after all steps issued so far) and starts an asynchronous
 * device->pinned-host copy on a second stream; it returns immediately and the copy overlaps the following nbx_step calls.
 * At most two snapshots can be in flight. nbx_stream_positions_end waits for the OLDEST snapshot in flight and copies it
 * (packed vec<T,N>[n], the layout of positions.bin frames) to x_host. */
int nbx_stream_positions_begin(nbx_engine* e);
int nbx_stream_positions_end(nbx_engine* e, void* x_host);

/* ---- multi-GPU plumbing (one process per GPU; rendezvous is the caller's, e.g. torch.distributed) ---------- */
#define NBX_UNIQUE_ID_BYTES 128
int nbx_comm_unique_id(void* id128);
int nbx_comm_init_rank(nbx_engine* e, const void* id128);

/* Peer-memory exchange for the BVH walk (one process per GPU on one NVLink/NVSwitch box). By default a rank's
 * accelerations reach the other ranks with an ncclAllGather after the walk. With peer buffers the walk kernel itself
 * stores every finished body's acceleration into ALL ranks' arrays (P2P stores over NVLink, overlapped with the walk
 * warp by warp) and only a barrier follows. nbx_peer_export writes this engine's two CUDA IPC memory handles (the
 * acceleration array alternates between two buffers, src/bvh.h:71-91 permutes every step); the caller gathers the
 * NBX_PEER_HANDLE_BYTES of every rank (rank order) and passes them to nbx_peer_import, which fails with NBX_ERR_COMM —
 * leaving the NCCL path in place — where CUDA IPC is not available; nbx_peer_import(e, NULL) drops the peer buffers again.
 * EVERY rank must end up on the same path (the caller agrees on it, e.g. with an all-reduce of the return codes).
 * Needs nbx_comm_init_rank first; NBX_BVH engines. */
#define NBX_PEER_HANDLE_BYTES 128
int nbx_peer_export(nbx_engine* e, void* handle128);
int nbx_peer_import(nbx_engine* e, const void* handles_world_x_128);

/* ---- measurement helpers ---------------------------------------------------------------------------------- */
/* FMA-pipe peak microbenchmark on the engine's device: precision NBX_F32 -> FFMA, NBX_F64 -> DFMA.
 * *tflops = 2 * fma/s / 1e12 measured with CUDA events. */
int nbx_measure_fma_peak(int device, int precision, double* tflops);
/* Tree engines: re-runs the traversal of the LAST built tree in counting mode (no state is modified) and returns, for
 * this rank's targets, the number of (body, node) tests, of accepted interactions (body-level pairs count as one), and
 * of warp-level steps (= records loaded per warp). Used for the issue-rate roofline of the walk (instructions per warp
 * step x warp_steps / walk time, bench.py) and for the parity check that the device performs exactly the oracle's tests. */
int nbx_traversal_stats(nbx_engine* e, uint64_t* node_visits, uint64_t* interactions, uint64_t* warp_steps);
/* bodies a warp of the tree walk serves per step: lanes carrying bodies (32; 16 or 8 for small problems, whose walks end
 * with their longest warp) x bodies per lane (octree: 1; bvh: 1, 2 or 4). The denominator of the lane utilisation
 * node_visits / (warp_steps x width). */
int nbx_walk_width(nbx_engine* e, uint32_t* bodies_per_warp_step);
/* counters of the engine since creation: kernels launched by this library, bytes copied H2D / D2H */
int nbx_get_counters(nbx_engine* e, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* per-phase device time (ms) of the last nbx_step* call's LAST step;
 * slot order: 0 force (all-pairs), 1 accel, 2 bbox/bounds, 3 sort (+keys), 4 build (octree cells), 5 multipoles
 * (bvh build_tree / octree monopoles), 6 traverse, 7 comm. Needs nbx_set_phase_timing(e,1). */
int nbx_set_phase_timing(nbx_engine* e, int enable);
int nbx_get_phase_ms(nbx_engine* e, float* ms, int capacity, int* count);

#ifdef __cplusplus
}
#endif
#endif /* NBX_H */
