// nbx_comm.cu — multi-GPU exchange of accelerations (one process per GPU, NCCL over NVLink 5 / NVSwitch).
//
// The reference has no distributed code; this is new work. Every rank keeps the whole state_t (replicated);
// the FORCE work is sharded and the accelerations are exchanged each step:
//   * ordered all-pairs / collapsed / bvh / octree: targets sharded by rank (rank r owns [r*chunk, (r+1)*chunk)), the
//     shards of `a` are all-gathered IN PLACE, then every rank integrates all bodies;
//   * symmetric all-pairs: (I, J) block-pair units dealt round-robin, per-rank sums combined with one all-reduce.
// NCCL is loaded with dlopen so that the single-GPU library has no link-time dependency (inside a torch process this
// resolves to torch's bundled libnccl.so.2, in the C++ driver to the system one).
#include <dlfcn.h>

#include <cstring>

#include <mutex>

#include "nbx_internal.cuh"

namespace nbx {

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int ncclChar = 0, ncclInt = 2, ncclFloat = 7, ncclDouble = 8, ncclSum = 0;

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                                           = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                                     = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)                                                               = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t)                  = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t)             = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t)             = nullptr;
  ncclResult_t (*GroupStart)()                                                                          = nullptr;
  ncclResult_t (*GroupEnd)()                                                                            = nullptr;
  const char* (*GetErrorString)(ncclResult_t)                                                           = nullptr;
  bool ok = false;
};

static void load_api(NcclApi& a) {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) return;
  a.GetUniqueId    = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
  a.CommInitRank   = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
  a.CommDestroy    = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
  a.AllGather      = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
  a.AllReduce      = (decltype(a.AllReduce))dlsym(a.lib, "ncclAllReduce");
  a.Broadcast      = (decltype(a.Broadcast))dlsym(a.lib, "ncclBroadcast");
  a.GroupStart     = (decltype(a.GroupStart))dlsym(a.lib, "ncclGroupStart");
  a.GroupEnd       = (decltype(a.GroupEnd))dlsym(a.lib, "ncclGroupEnd");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
  a.ok             = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce && a.Broadcast && a.GroupStart && a.GroupEnd && a.GetErrorString;
}

// the C++ driver calls in from one host thread per GPU: the table is filled exactly once, before anyone reads it
NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, load_api, std::ref(a));
  return a;
}

int nccl_fail(const char* what, ncclResult_t r) {
  return fail(NBX_ERR_COMM, std::string(what) + ": " + (api().GetErrorString ? api().GetErrorString(r) : "nccl error"));
}
}  // namespace

int comm_unique_id(void* id128) {
  if (!id128) return fail(NBX_ERR_INVALID, "id128 is NULL");
  if (!api().ok) return fail(NBX_ERR_COMM, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  ncclResult_t r = api().GetUniqueId(&id);
  if (r != 0) return nccl_fail("ncclGetUniqueId", r);
  memcpy(id128, &id, sizeof(id));
  return NBX_OK;
}

int comm_init_rank(nbx_engine* e, const void* id128) {
  if (!id128) return fail(NBX_ERR_INVALID, "id128 is NULL");
  if (e->cfg.world_size <= 1) return NBX_OK;
  if (!api().ok) return fail(NBX_ERR_COMM, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c   = nullptr;
  ncclResult_t r = api().CommInitRank(&c, e->cfg.world_size, id, e->cfg.rank);
  if (r != 0) return nccl_fail("ncclCommInitRank", r);
  e->comm = c;
  return NBX_OK;
}

// in-place all-gather of one vec4 array: rank r contributes records [r*chunk, (r+1)*chunk)
int comm_allgather(nbx_engine* e, void* vec4_array) {
  if (e->cfg.world_size <= 1) return NBX_OK;
  if (!e->comm) return fail(NBX_ERR_COMM, "multi-GPU engine used before nbx_comm_init_rank");
  PhaseTimer pt(e, PH_COMM);
  char* base         = static_cast<char*>(vec4_array);
  const size_t bytes = rec_bytes(e) * e->chunk;
  ncclResult_t r = api().AllGather(base + bytes * e->cfg.rank, base, bytes, ncclChar, (ncclComm_t)e->comm, e->stream);
  if (r != 0) return nccl_fail("ncclAllGather", r);
  return NBX_OK;
}

// sum over ranks, in place (symmetric all-pairs: per-rank partial accelerations -> total)
int comm_allreduce_sum(nbx_engine* e, void* buffer, size_t count) {
  if (e->cfg.world_size <= 1) return NBX_OK;
  if (!e->comm) return fail(NBX_ERR_COMM, "multi-GPU engine used before nbx_comm_init_rank");
  PhaseTimer pt(e, PH_COMM);
  ncclResult_t r = api().AllReduce(buffer, buffer, count, e->prec == 4 ? ncclFloat : ncclDouble, ncclSum, (ncclComm_t)e->comm, e->stream);
  if (r != 0) return nccl_fail("ncclAllReduce", r);
  return NBX_OK;
}

// in-place broadcast of `bytes` bytes from rank `root` (build-once-and-broadcast variant of the tree build)
int comm_broadcast(nbx_engine* e, void* buffer, size_t bytes, int root) {
  if (e->cfg.world_size <= 1) return NBX_OK;
  if (!e->comm) return fail(NBX_ERR_COMM, "multi-GPU engine used before nbx_comm_init_rank");
  ncclResult_t r = api().Broadcast(buffer, buffer, bytes, ncclChar, root, (ncclComm_t)e->comm, e->stream);
  if (r != 0) return nccl_fail("ncclBroadcast", r);
  return NBX_OK;
}

// all-gather with per-rank sizes: one grouped broadcast per owner
int comm_allgatherv(nbx_engine* e, void* buffer, const size_t* offset, const size_t* count) {
  if (e->cfg.world_size <= 1) return NBX_OK;
  if (!e->comm) return fail(NBX_ERR_COMM, "multi-GPU engine used before nbx_comm_init_rank");
  PhaseTimer pt(e, PH_COMM);
  char* base     = static_cast<char*>(buffer);
  ncclResult_t r = api().GroupStart();
  for (int q = 0; q < e->cfg.world_size && r == 0; ++q)
    if (count[q]) r = api().Broadcast(base + offset[q], base + offset[q], count[q], ncclChar, q, (ncclComm_t)e->comm, e->stream);
  const ncclResult_t r2 = api().GroupEnd();
  if (r != 0 || r2 != 0) return nccl_fail("grouped ncclBroadcast", r != 0 ? r : r2);
  return NBX_OK;
}

// stream-ordered barrier over the ranks: a 4-byte all-reduce
int comm_barrier(nbx_engine* e) {
  if (e->cfg.world_size <= 1) return NBX_OK;
  if (!e->comm) return fail(NBX_ERR_COMM, "multi-GPU engine used before nbx_comm_init_rank");
  PhaseTimer pt(e, PH_COMM);
  if (!e->barrier_scratch) {
    NBX_CUDA(cudaMalloc(&e->barrier_scratch, sizeof(int)));
    NBX_CUDA(cudaMemsetAsync(e->barrier_scratch, 0, sizeof(int), e->stream));
  }
  ncclResult_t r = api().AllReduce(e->barrier_scratch, e->barrier_scratch, 1, ncclInt, ncclSum, (ncclComm_t)e->comm, e->stream);
  if (r != 0) return nccl_fail("ncclAllReduce (barrier)", r);
  return NBX_OK;
}

// ---- peer buffers (CUDA IPC) --------------------------------------------------------------------------------------------
int peer_export(nbx_engine* e, void* handle128) {
  if (!handle128) return fail(NBX_ERR_INVALID, "handle buffer is NULL");
  if (e->algo != NBX_BVH || !e->own_a[0] || !e->own_a[1]) return fail(NBX_ERR_STATE, "peer buffers exist for NBX_BVH engines only");
  static_assert(2 * sizeof(cudaIpcMemHandle_t) == NBX_PEER_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h[2];
  for (int k = 0; k < 2; ++k) {
    cudaError_t err = cudaIpcGetMemHandle(&h[k], e->own_a[k]);
    if (err != cudaSuccess) {
      cudaGetLastError();
      return fail(NBX_ERR_COMM, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(err));
    }
  }
  memcpy(handle128, h, sizeof(h));
  return NBX_OK;
}

int peer_import(nbx_engine* e, const void* handles) {
  if (!handles) {  // back to the NCCL all-gather (every rank must take the same path: the caller agrees on it)
    peer_close(e);
    return NBX_OK;
  }
  const int world = e->cfg.world_size, rank = e->cfg.rank;
  if (world < 2 || world > 16) return fail(NBX_ERR_INVALID, "peer buffers need 2..16 ranks");
  if (e->algo != NBX_BVH || !e->own_a[0]) return fail(NBX_ERR_STATE, "peer buffers exist for NBX_BVH engines only");
  if (!e->comm) return fail(NBX_ERR_COMM, "nbx_peer_import before nbx_comm_init_rank");
  peer_close(e);
  const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handles);
  for (int r = 0; r < world; ++r)
    for (int k = 0; k < 2; ++k) {
      if (r == rank) { e->peer_a[k][r] = e->own_a[k]; continue; }
      cudaIpcMemHandle_t hh;
      memcpy(&hh, &h[2 * r + k], sizeof(hh));
      void* p = nullptr;
      cudaError_t err = cudaIpcOpenMemHandle(&p, hh, cudaIpcMemLazyEnablePeerAccess);
      if (err != cudaSuccess) {
        cudaGetLastError();
        peer_close(e);
        return fail(NBX_ERR_COMM, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(err) + " (the NCCL all-gather stays in use)");
      }
      e->peer_a[k][r] = p;
    }
  e->peers_ready = true;
  return NBX_OK;
}

void peer_close(nbx_engine* e) {
  for (int k = 0; k < 2; ++k)
    for (int r = 0; r < 16; ++r) {
      if (e->peer_a[k][r] && e->peer_a[k][r] != e->own_a[k]) cudaIpcCloseMemHandle(e->peer_a[k][r]);
      e->peer_a[k][r] = nullptr;
    }
  e->peers_ready = false;
}

void comm_destroy(nbx_engine* e) {
  peer_close(e);
  if (e->barrier_scratch) cudaFree(e->barrier_scratch);
  e->barrier_scratch = nullptr;
  if (e->comm && api().ok) api().CommDestroy((ncclComm_t)e->comm);
  e->comm = nullptr;
}

}  // namespace nbx
