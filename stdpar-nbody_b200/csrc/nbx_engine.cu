// nbx_engine.cu — engine lifetime, host<->device state transfer, step drivers and the extern "C" surface of nbx.h.
#include <cstdlib>
#include <cstring>

#include "nbx_internal.cuh"

namespace nbx {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

// ---- AoS (state_t, src/system.h:41-50) <-> vec4 records -------------------------------------------------------
// pack: stage holds T m[n] | T q[n*D]  ->  xm[i] = (x, y, z|0, m)   or   vec[i] = (q0, q1, q2|0, 0)
template <typename T, int D>
__global__ void pack_pos_kernel(const T* __restrict__ x, vec4_t<T>* xm, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  vec4_t<T> r = xm[i];
  r.x = x[size_t(i) * D];
  r.y = x[size_t(i) * D + 1];
  r.z = D == 3 ? x[size_t(i) * D + 2] : T(0);
  xm[i] = r;
}
template <typename T>
__global__ void pack_mass_kernel(const T* __restrict__ m, vec4_t<T>* xm, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  vec4_t<T> r = xm[i];
  r.w = m[i];
  xm[i] = r;
}
template <typename T, int D>
__global__ void pack_vec_kernel(const T* __restrict__ q, vec4_t<T>* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = make_v4<T>(q[size_t(i) * D], q[size_t(i) * D + 1], D == 3 ? q[size_t(i) * D + 2] : T(0), T(0));
}
template <typename T, int D>
__global__ void unpack_vec_kernel(const vec4_t<T>* __restrict__ in, T* q, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  vec4_t<T> r = in[i];
  q[size_t(i) * D]     = r.x;
  q[size_t(i) * D + 1] = r.y;
  if (D == 3) q[size_t(i) * D + 2] = r.z;
}
template <typename T>
__global__ void unpack_mass_kernel(const vec4_t<T>* __restrict__ in, T* m, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  m[i] = in[i].w;
}

static int ensure_stage(nbx_engine* e, size_t bytes) {
  if (bytes <= e->stage_bytes) return NBX_OK;
  if (e->stage) cudaFree(e->stage);
  e->stage = nullptr;
  NBX_CUDA(cudaMalloc(&e->stage, bytes));
  e->stage_bytes = bytes;
  return NBX_OK;
}

// [first, first + count) of the bodies; host pointers address the FULL arrays. gather: all-gather the uploaded shards
// over NCCL afterwards (nbx_upload_shard: every rank sends 1/world of the bytes over PCIe, NVLink does the rest).
template <typename T, int D>
static int upload_impl(nbx_engine* e, const void* m, const void* x, const void* v, const void* a, const void* ao, uint32_t first,
                       uint32_t count, bool gather) {
  const dim3 grid((count + 255) / 256), block(256);
  NBX_TRY(ensure_stage(e, sizeof(T) * size_t(e->n) * D));
  auto put = [&](const void* host, size_t elems_per_body) -> int {
    const size_t bytes = sizeof(T) * elems_per_body * count;
    if (!bytes) return NBX_OK;
    NBX_CUDA(cudaMemcpyAsync(e->stage, static_cast<const T*>(host) + elems_per_body * first, bytes, cudaMemcpyHostToDevice, e->stream));
    e->h2d += bytes;
    return NBX_OK;
  };
  vec4_t<T>* xm = static_cast<vec4_t<T>*>(e->xm[e->cur]) + first;
  if (x) {
    NBX_TRY(put(x, D));
    if (count) pack_pos_kernel<T, D><<<grid, block, 0, e->stream>>>((const T*)e->stage, xm, count);
    e->launches++;
  }
  if (m) {
    NBX_TRY(put(m, 1));
    if (count) pack_mass_kernel<T><<<grid, block, 0, e->stream>>>((const T*)e->stage, xm, count);
    e->launches++;
  }
  if (x || m) {
    if (gather) NBX_TRY(comm_allgather(e, e->xm[e->cur]));
    // both position buffers carry the masses (the fused leapfrog writes x,m into the other one)
    NBX_CUDA(cudaMemcpyAsync(e->xm[e->cur ^ 1], e->xm[e->cur], rec_bytes(e) * e->n, cudaMemcpyDeviceToDevice, e->stream));
  }
  const void* src[3] = {v, a, ao};
  void* dst[3]       = {e->v, e->a, e->ao};
  for (int q = 0; q < 3; ++q) {
    if (!src[q]) continue;
    NBX_TRY(put(src[q], D));
    if (count) pack_vec_kernel<T, D><<<grid, block, 0, e->stream>>>((const T*)e->stage, static_cast<vec4_t<T>*>(dst[q]) + first, count);
    e->launches++;
    if (gather) NBX_TRY(comm_allgather(e, dst[q]));
  }
  NBX_CUDA(cudaGetLastError());
  NBX_CUDA(cudaStreamSynchronize(e->stream));  // host arrays may be reused by the caller
  return NBX_OK;
}

template <typename T, int D>
static int download_impl(nbx_engine* e, void* m, void* x, void* v, void* a, void* ao, uint32_t first, uint32_t count) {
  const dim3 grid((count + 255) / 256), block(256);
  NBX_TRY(ensure_stage(e, sizeof(T) * size_t(e->n) * D));
  if (!count) return NBX_OK;
  auto get = [&](void* host, size_t elems_per_body) -> int {
    const size_t bytes = sizeof(T) * elems_per_body * count;
    NBX_CUDA(cudaMemcpyAsync(static_cast<T*>(host) + elems_per_body * first, e->stage, bytes, cudaMemcpyDeviceToHost, e->stream));
    NBX_CUDA(cudaStreamSynchronize(e->stream));
    e->d2h += bytes;
    return NBX_OK;
  };
  if (m) {
    unpack_mass_kernel<T><<<grid, block, 0, e->stream>>>(static_cast<const vec4_t<T>*>(e->xm[e->cur]) + first, (T*)e->stage, count);
    e->launches++;
    NBX_TRY(get(m, 1));
  }
  void* dsth[4]      = {x, v, a, ao};
  const void* src[4] = {e->xm[e->cur], e->v, e->a, e->ao};
  for (int q = 0; q < 4; ++q) {
    if (!dsth[q]) continue;
    unpack_vec_kernel<T, D><<<grid, block, 0, e->stream>>>(static_cast<const vec4_t<T>*>(src[q]) + first, (T*)e->stage, count);
    e->launches++;
    NBX_TRY(get(dsth[q], D));
  }
  NBX_CUDA(cudaGetLastError());
  return NBX_OK;
}

// ---- Saver streaming ------------------------------------------------------------------------------------------------
template <typename T, int D>
static int stream_begin_impl(nbx_engine* e) {
  if (e->snap_count == 2) return fail(NBX_ERR_STATE, "two position snapshots already in flight: call nbx_stream_positions_end first");
  const size_t bytes = sizeof(T) * size_t(e->n) * D;
  if (!e->copy_stream) {
    NBX_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      NBX_CUDA(cudaMalloc(&e->snap_dev[k], bytes));
      NBX_CUDA(cudaMallocHost(&e->snap_host[k], bytes));
      NBX_CUDA(cudaEventCreateWithFlags(&e->snap_ready[k], cudaEventDisableTiming));
      NBX_CUDA(cudaEventCreateWithFlags(&e->snap_done[k], cudaEventDisableTiming));
    }
  }
  const int slot = (e->snap_head + e->snap_count) % 2;
  // the (cheap, HBM-bound) unpack runs on the main stream, so later steps may overwrite the positions freely; the slow
  // PCIe copy runs on the copy stream and overlaps them
  unpack_vec_kernel<T, D><<<(e->n + 255) / 256, 256, 0, e->stream>>>((const vec4_t<T>*)e->xm[e->cur], (T*)e->snap_dev[slot], e->n);
  e->launches++;
  NBX_CUDA(cudaEventRecord(e->snap_ready[slot], e->stream));
  NBX_CUDA(cudaStreamWaitEvent(e->copy_stream, e->snap_ready[slot], 0));
  NBX_CUDA(cudaMemcpyAsync(e->snap_host[slot], e->snap_dev[slot], bytes, cudaMemcpyDeviceToHost, e->copy_stream));
  NBX_CUDA(cudaEventRecord(e->snap_done[slot], e->copy_stream));
  e->d2h += bytes;
  e->snap_count++;
  return NBX_OK;
}

template <typename T, int D>
static int stream_end_impl(nbx_engine* e, void* x_host) {
  if (e->snap_count == 0) return fail(NBX_ERR_STATE, "no position snapshot in flight");
  if (!x_host) return fail(NBX_ERR_INVALID, "x_host is NULL");
  const int slot = e->snap_head;
  NBX_CUDA(cudaEventSynchronize(e->snap_done[slot]));
  memcpy(x_host, e->snap_host[slot], sizeof(T) * size_t(e->n) * D);
  e->snap_head = (e->snap_head + 1) % 2;
  e->snap_count--;
  return NBX_OK;
}

#define NBX_DISPATCH(e, fn, ...)                                                          \
  ((e)->prec == 4 ? ((e)->dim == 2 ? fn<float, 2>(__VA_ARGS__) : fn<float, 3>(__VA_ARGS__)) \
                  : ((e)->dim == 2 ? fn<double, 2>(__VA_ARGS__) : fn<double, 3>(__VA_ARGS__)))

// ---- one time step (the `kernels()` lambda of the reference drivers) -------------------------------------------
static int bvh_step_enqueue(nbx_engine* e) {
  NBX_TRY(bvh_bounding_box(e));
  NBX_TRY(bvh_hilbert_sort(e));
  NBX_TRY(bvh_build_tree(e));
  NBX_TRY(bvh_compute_force(e));  // all-gathers a[tb,te) of every rank -> full a everywhere
  return accelerate_step(e);
}

// The BVH step is ~30 launches (bbox 2, keys, sort 11, gather, one per build level, walk, leapfrog) with no host decision in
// between: on one GPU it is captured ONCE per buffer parity into a CUDA graph and replayed, which removes the per-launch
// host work and most of the gaps between the small kernels (what small n is bound by). NBX_GRAPH=0 disables it; phase
// timing (events between the kernels) and multi-GPU steps (NCCL calls) use the plain launch sequence.
static int bvh_step(nbx_engine* e) {
  if (!e->use_graph || e->phase_timing || e->cfg.world_size > 1) return bvh_step_enqueue(e);
  const int par = e->cur;
  if (e->step_graph[par] && e->step_graph_v[par] == e->v) {
    NBX_CUDA(cudaGraphLaunch(e->step_graph[par], e->stream));
    bvh_after_graph_replay(e);
    e->launches += e->step_graph_launches[par];
    return NBX_OK;
  }
  if (e->step_graph[par]) {  // captured with another buffer assignment (per-phase calls in between): recapture
    cudaGraphExecDestroy(e->step_graph[par]);
    e->step_graph[par] = nullptr;
  }
  const void* v_before    = e->v;
  const uint64_t l_before = e->launches;
  NBX_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = bvh_step_enqueue(e);
  cudaGraph_t graph = nullptr;
  const cudaError_t err = cudaStreamEndCapture(e->stream, &graph);
  if (rc != NBX_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (err != cudaSuccess) return fail(NBX_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(err));
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ierr = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ierr != cudaSuccess) return fail(NBX_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ierr));
  e->step_graph[par]          = exec;
  e->step_graph_v[par]        = v_before;
  e->step_graph_launches[par] = e->launches - l_before;
  NBX_CUDA(cudaGraphLaunch(exec, e->stream));  // the capture only recorded the step: run it
  return NBX_OK;
}

static int one_step(nbx_engine* e) {
  const bool multi = e->cfg.world_size > 1;
  switch (e->algo) {
    case NBX_ALL_PAIRS: {
      // multi-GPU: the force functions leave the FULL acceleration array on every rank (symmetric kernel: all-reduce of
      // the per-rank sums; ordered kernel: all-gather of the per-rank target shards), then every rank integrates all
      // bodies — the whole state_t stays replicated and any rank can serve nbx_download / nbx_calc_energies.
      const bool fuse = !(e->cfg.flags & NBX_FLAG_NO_FUSED_INTEGRATE) && (!multi || all_pairs_sym_enabled(e));
      NBX_TRY(all_pairs_force(e, fuse));
      if (!fuse) NBX_TRY(accelerate_step(e));
      return NBX_OK;
    }
    case NBX_ALL_PAIRS_COLLAPSED:
      NBX_TRY(all_pairs_collapsed_force(e));
      NBX_TRY(accelerate_step(e));
      return NBX_OK;
    case NBX_BVH: return bvh_step(e);
    case NBX_OCTREE:
      NBX_TRY(octree_build(e));
      NBX_TRY(octree_compute_force(e));  // all-gathers the sorted-slot accelerations itself
      NBX_TRY(accelerate_step(e));
      return NBX_OK;
    default: return fail(NBX_ERR_INVALID, "unknown algorithm");
  }
}

static void collect_phase_times(nbx_engine* e) {
  if (!e->phase_timing) return;
  for (int s = 0; s < PH_COUNT; ++s) {
    e->ph_ms[s] = 0;
    if (e->ph_used[s]) cudaEventElapsedTime(&e->ph_ms[s], e->ph_ev[s][0], e->ph_ev[s][1]);
  }
}

}  // namespace nbx

using namespace nbx;

// =================================================================================================================
extern "C" {

const char* nbx_last_error(void) { return g_last_error.c_str(); }
int nbx_version(void) { return NBX_VERSION; }

int nbx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int nbx_create(const nbx_config* cfg, nbx_engine** out) {
  if (!cfg || !out) return fail(NBX_ERR_INVALID, "nbx_create: NULL argument");
  *out = nullptr;
  if (cfg->struct_size != sizeof(nbx_config)) return fail(NBX_ERR_INVALID, "nbx_create: struct_size mismatch");
  if (cfg->dim != 2 && cfg->dim != 3)
    return fail(NBX_ERR_INVALID, "nbx_create: dim must be 2 or 3 (reference: -DDIM_SIZE, src/main.cpp:5-7)");
  if (cfg->precision != NBX_F32 && cfg->precision != NBX_F64)
    return fail(NBX_ERR_INVALID, "nbx_create: precision must be NBX_F32 or NBX_F64");
  if (cfg->algorithm < NBX_ALL_PAIRS || cfg->algorithm > NBX_BVH) return fail(NBX_ERR_INVALID, "nbx_create: unknown algorithm");
  if (cfg->n == 0) return fail(NBX_ERR_INVALID, "nbx_create: n must be > 0");
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size)
    return fail(NBX_ERR_INVALID, "nbx_create: bad rank/world_size");
  if (cfg->algorithm == NBX_BVH && cfg->n < 2) return fail(NBX_ERR_INVALID, "nbx_create: bvh needs n >= 2");
  int ndev = nbx_device_count();
  if (ndev <= 0) return fail(NBX_ERR_NO_DEVICE, "nbx_create: no CUDA device (libnbx has no CPU path)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(NBX_ERR_INVALID, "nbx_create: device ordinal out of range");

  nbx_engine* e = new nbx_engine();
  e->cfg  = *cfg;
  e->dim  = cfg->dim;
  e->prec = cfg->precision;
  e->algo = cfg->algorithm;
  e->n    = cfg->n;
  e->device = cfg->device;
  {
    const char* v = getenv("NBX_GRAPH");
    e->use_graph  = !(v && atoi(v) == 0);
  }
  e->chunk  = (cfg->n + cfg->world_size - 1) / cfg->world_size;
  e->n_pad  = e->chunk * cfg->world_size;
  e->tb     = std::min<uint64_t>(uint64_t(e->chunk) * cfg->rank, cfg->n);
  e->te     = std::min<uint64_t>(uint64_t(e->chunk) * (cfg->rank + 1), cfg->n);
  auto bail = [&](int rc) { nbx_destroy(e); return rc; };
#define NBX_CUDA_B(call)                                                                          \
  do {                                                                                            \
    cudaError_t err__ = (call);                                                                   \
    if (err__ != cudaSuccess) return bail(fail(NBX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__))); \
  } while (0)
  NBX_CUDA_B(cudaSetDevice(e->device));
  cudaDeviceProp prop;
  NBX_CUDA_B(cudaGetDeviceProperties(&prop, e->device));
  e->sm_count = prop.multiProcessorCount;
  NBX_CUDA_B(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  NBX_CUDA_B(cudaEventCreate(&e->ev0));
  NBX_CUDA_B(cudaEventCreate(&e->ev1));
  for (int s = 0; s < PH_COUNT; ++s)
    for (int k = 0; k < 2; ++k) NBX_CUDA_B(cudaEventCreate(&e->ph_ev[s][k]));
  const size_t rb = rec_bytes(e);
  // position buffers carry a zero-mass tail (n_pad - n records of padding + one all-pairs tile) so that tile
  // loads never need a bounds check: a zero-mass source contributes exactly 0.
  size_t pos_records = size_t(e->n_pad) + 1024;
  if (e->algo == NBX_ALL_PAIRS || e->algo == NBX_ALL_PAIRS_COLLAPSED) {  // the symmetric kernel reads whole blocks of B bodies
    const size_t B = all_pairs_sym_block(e->n, cfg->world_size);
    pos_records    = std::max(pos_records, (size_t(e->n) + B - 1) / B * B + 1024);
  }
  for (int k = 0; k < 2; ++k) {
    NBX_CUDA_B(cudaMalloc(&e->xm[k], rb * pos_records));
    NBX_CUDA_B(cudaMemsetAsync(e->xm[k], 0, rb * pos_records, e->stream));
  }
  void** vecs[3] = {&e->v, &e->a, &e->ao};
  for (auto pp : vecs) {
    NBX_CUDA_B(cudaMalloc(pp, rb * e->n_pad));
    NBX_CUDA_B(cudaMemsetAsync(*pp, 0, rb * e->n_pad, e->stream));
  }
  int rc = NBX_OK;
  if (e->algo == NBX_BVH) rc = bvh_create(e);
  if (e->algo == NBX_OCTREE) rc = octree_create(e);
  if (rc != NBX_OK) return bail(rc);
  NBX_CUDA_B(cudaStreamSynchronize(e->stream));
#undef NBX_CUDA_B
  *out = e;
  return NBX_OK;
}

int nbx_destroy(nbx_engine* e) {
  if (!e) return NBX_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (auto& g : e->step_graph)
    if (g) cudaGraphExecDestroy(g);
  comm_destroy(e);
  bvh_destroy(e);
  octree_destroy(e);
  sorter_destroy(e);
  all_pairs_sym_destroy(e);
  void* bufs[] = {e->xm[0], e->xm[1], e->v, e->a, e->ao, e->v_alt, e->a_alt, e->ao_alt, e->partial, e->tickets, e->stage, e->energy_out};
  for (void* b : bufs)
    if (b) cudaFree(b);
  for (int k = 0; k < 2; ++k) {
    if (e->snap_dev[k]) cudaFree(e->snap_dev[k]);
    if (e->snap_host[k]) cudaFreeHost(e->snap_host[k]);
    if (e->snap_ready[k]) cudaEventDestroy(e->snap_ready[k]);
    if (e->snap_done[k]) cudaEventDestroy(e->snap_done[k]);
  }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  for (int s = 0; s < PH_COUNT; ++s)
    for (int k = 0; k < 2; ++k)
      if (e->ph_ev[s][k]) cudaEventDestroy(e->ph_ev[s][k]);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return NBX_OK;
}

#define NBX_ENTER(e)                                                   \
  if (!(e)) return fail(NBX_ERR_INVALID, std::string(__func__) + ": NULL engine"); \
  NBX_CUDA(cudaSetDevice((e)->device))

int nbx_upload(nbx_engine* e, const void* m, const void* x, const void* v, const void* a, const void* ao) {
  NBX_ENTER(e);
  return NBX_DISPATCH(e, upload_impl, e, m, x, v, a, ao, 0u, e->n, false);
}

int nbx_download(nbx_engine* e, void* m, void* x, void* v, void* a, void* ao) {
  NBX_ENTER(e);
  return NBX_DISPATCH(e, download_impl, e, m, x, v, a, ao, 0u, e->n);
}

int nbx_upload_shard(nbx_engine* e, const void* m, const void* x, const void* v, const void* a, const void* ao) {
  NBX_ENTER(e);
  if (e->cfg.world_size > 1 && !e->comm) return fail(NBX_ERR_COMM, "nbx_upload_shard before nbx_comm_init_rank");
  return NBX_DISPATCH(e, upload_impl, e, m, x, v, a, ao, e->tb, e->te - e->tb, e->cfg.world_size > 1);
}

int nbx_download_shard(nbx_engine* e, void* m, void* x, void* v, void* a, void* ao) {
  NBX_ENTER(e);
  return NBX_DISPATCH(e, download_impl, e, m, x, v, a, ao, e->tb, e->te - e->tb);
}

int nbx_stream_positions_begin(nbx_engine* e) {
  NBX_ENTER(e);
  return NBX_DISPATCH(e, stream_begin_impl, e);
}
int nbx_stream_positions_end(nbx_engine* e, void* x_host) {
  NBX_ENTER(e);
  return NBX_DISPATCH(e, stream_end_impl, e, x_host);
}

int nbx_step(nbx_engine* e, uint32_t steps) {
  NBX_ENTER(e);
  for (uint32_t s = 0; s < steps; ++s) NBX_TRY(one_step(e));
  return NBX_OK;
}

int nbx_step_timed(nbx_engine* e, uint32_t steps, float* ms) {
  NBX_ENTER(e);
  NBX_CUDA(cudaEventRecord(e->ev0, e->stream));
  for (uint32_t s = 0; s < steps; ++s) NBX_TRY(one_step(e));
  NBX_CUDA(cudaEventRecord(e->ev1, e->stream));
  NBX_CUDA(cudaEventSynchronize(e->ev1));
  if (ms) NBX_CUDA(cudaEventElapsedTime(ms, e->ev0, e->ev1));
  collect_phase_times(e);
  return NBX_OK;
}

int nbx_sync(nbx_engine* e) {
  NBX_ENTER(e);
  NBX_CUDA(cudaStreamSynchronize(e->stream));
  collect_phase_times(e);
  return NBX_OK;
}

int nbx_all_pairs_force(nbx_engine* e) {
  NBX_ENTER(e);
  return all_pairs_force(e, false);
}
int nbx_all_pairs_collapsed_force(nbx_engine* e) {
  NBX_ENTER(e);
  return all_pairs_collapsed_force(e);
}
int nbx_accelerate_step(nbx_engine* e) {
  NBX_ENTER(e);
  return accelerate_step(e);
}
int nbx_calc_energies(nbx_engine* e, double* kinetic, double* gravitational) {
  NBX_ENTER(e);
  return calc_energies(e, kinetic, gravitational);
}

int nbx_bvh_bounding_box(nbx_engine* e, void* xmin, void* xmax) {
  NBX_ENTER(e);
  if (e->algo != NBX_BVH) return fail(NBX_ERR_STATE, "engine was not created with NBX_BVH");
  NBX_TRY(bvh_bounding_box(e));
  if (xmin || xmax) return bvh_get_bbox(e, xmin, xmax);
  return NBX_OK;
}
int nbx_bvh_hilbert_sort(nbx_engine* e) {
  NBX_ENTER(e);
  if (e->algo != NBX_BVH) return fail(NBX_ERR_STATE, "engine was not created with NBX_BVH");
  return bvh_hilbert_sort(e);
}
int nbx_bvh_build_tree(nbx_engine* e) {
  NBX_ENTER(e);
  if (e->algo != NBX_BVH) return fail(NBX_ERR_STATE, "engine was not created with NBX_BVH");
  return bvh_build_tree(e);
}
int nbx_bvh_compute_force(nbx_engine* e) {
  NBX_ENTER(e);
  if (e->algo != NBX_BVH) return fail(NBX_ERR_STATE, "engine was not created with NBX_BVH");
  return bvh_compute_force(e);
}
int nbx_bvh_get_keys(nbx_engine* e, uint64_t* keys, uint32_t* perm) {
  NBX_ENTER(e);
  if (e->algo != NBX_BVH) return fail(NBX_ERR_STATE, "engine was not created with NBX_BVH");
  return bvh_get_keys(e, keys, perm);
}
int nbx_bvh_get_nodes(nbx_engine* e, uint64_t* nnodes, void* node_m, void* bw, void* b) {
  NBX_ENTER(e);
  if (e->algo != NBX_BVH) return fail(NBX_ERR_STATE, "engine was not created with NBX_BVH");
  return bvh_get_nodes(e, nnodes, node_m, bw, b);
}

int nbx_octree_build(nbx_engine* e) {
  NBX_ENTER(e);
  if (e->algo != NBX_OCTREE) return fail(NBX_ERR_STATE, "engine was not created with NBX_OCTREE");
  return octree_build(e);  // reports NBX_ERR_CAPACITY itself
}
int nbx_octree_compute_force(nbx_engine* e) {
  NBX_ENTER(e);
  if (e->algo != NBX_OCTREE) return fail(NBX_ERR_STATE, "engine was not created with NBX_OCTREE");
  return octree_compute_force(e);
}
int nbx_octree_get_root(nbx_engine* e, void* side, void* root_x, uint64_t* nodes_used) {
  NBX_ENTER(e);
  if (e->algo != NBX_OCTREE) return fail(NBX_ERR_STATE, "engine was not created with NBX_OCTREE");
  return octree_get_root(e, side, root_x, nodes_used);
}
int nbx_octree_get_canonical(nbx_engine* e, uint64_t* count, uint32_t* depth, uint64_t* path, uint32_t* kind,
                             void* monopole) {
  NBX_ENTER(e);
  if (e->algo != NBX_OCTREE) return fail(NBX_ERR_STATE, "engine was not created with NBX_OCTREE");
  return octree_get_canonical(e, count, depth, path, kind, monopole);
}

int nbx_traversal_stats(nbx_engine* e, uint64_t* node_visits, uint64_t* interactions, uint64_t* warp_steps) {
  NBX_ENTER(e);
  if (e->algo != NBX_BVH && e->algo != NBX_OCTREE) return fail(NBX_ERR_STATE, "traversal stats need a tree engine");
  unsigned long long* d = nullptr;
  NBX_CUDA(cudaMalloc(&d, 3 * sizeof(unsigned long long)));
  NBX_CUDA(cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), e->stream));
  int rc = e->algo == NBX_BVH ? bvh_stats(e, d) : octree_stats(e, d);
  unsigned long long h[3] = {0, 0, 0};
  if (rc == NBX_OK) {
    cudaError_t err = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) rc = fail(NBX_ERR_CUDA, cudaGetErrorString(err));
  }
  cudaFree(d);
  if (rc != NBX_OK) return rc;
  if (node_visits) *node_visits = h[0];
  if (interactions) *interactions = h[1];
  if (warp_steps) *warp_steps = h[2];
  return NBX_OK;
}

int nbx_walk_width(nbx_engine* e, uint32_t* bodies_per_warp_step) {
  if (!e || !bodies_per_warp_step) return fail(NBX_ERR_INVALID, "NULL argument");
  if (e->algo != NBX_BVH && e->algo != NBX_OCTREE) return fail(NBX_ERR_STATE, "walk width needs a tree engine");
  *bodies_per_warp_step = e->algo == NBX_BVH ? uint32_t(bvh_walk_width(e)) : (e->algo == NBX_OCTREE ? uint32_t(octree_walk_width(e)) : 32u);
  return NBX_OK;
}

int nbx_peer_export(nbx_engine* e, void* handle128) {
  NBX_ENTER(e);
  return peer_export(e, handle128);
}
int nbx_peer_import(nbx_engine* e, const void* handles) {
  NBX_ENTER(e);
  return peer_import(e, handles);
}

int nbx_comm_unique_id(void* id128) { return comm_unique_id(id128); }
int nbx_comm_init_rank(nbx_engine* e, const void* id128) {
  NBX_ENTER(e);
  return comm_init_rank(e, id128);
}

int nbx_measure_fma_peak(int device, int precision, double* tflops) {
  if (nbx_device_count() <= 0) return fail(NBX_ERR_NO_DEVICE, "no CUDA device");
  return measure_fma_peak(device, precision, tflops);
}

int nbx_get_counters(nbx_engine* e, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  if (!e) return fail(NBX_ERR_INVALID, "NULL engine");
  if (kernel_launches) *kernel_launches = e->launches;
  if (h2d_bytes) *h2d_bytes = e->h2d;
  if (d2h_bytes) *d2h_bytes = e->d2h;
  return NBX_OK;
}

int nbx_set_phase_timing(nbx_engine* e, int enable) {
  if (!e) return fail(NBX_ERR_INVALID, "NULL engine");
  e->phase_timing = enable != 0;
  for (int s = 0; s < PH_COUNT; ++s) e->ph_used[s] = false;
  return NBX_OK;
}

int nbx_get_phase_ms(nbx_engine* e, float* ms, int capacity, int* count) {
  if (!e || !ms) return fail(NBX_ERR_INVALID, "NULL argument");
  int k = 0;
  for (int s = 0; s < PH_COUNT && k < capacity; ++s) ms[k++] = e->ph_used[s] ? e->ph_ms[s] : 0.f;
  if (count) *count = k;
  return NBX_OK;
}

}  // extern "C"
