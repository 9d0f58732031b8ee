// The one collective of the path: an all-gather of per-object feature rows for well-level
// aggregation and normalisation (north_star; parity target Normalize_CP_ami.py:126,
// Pycyto_pertime.py:69-72 -- the reference itself has no collective).
//
// NCCL is bound at run time (dlopen of libnccl.so.2, which resolves to the copy PyTorch has
// already loaded when the process uses torch.distributed) so that libips.so itself has no
// link-time dependency on it.  The communicator is created from a unique id that the host
// side broadcasts over its existing control plane (torch.distributed in this repository).
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "ips_common.cuh"

namespace ips {

// Minimal NCCL ABI (nccl.h: stable since 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclInt64 = 4 };

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(h, "ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
  });
  return api;
}

struct Comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, world = 1;
  int64_t* d_counts = nullptr;   // [world] staging for the row-count exchange
};

#define IPS_NCCL_OK(expr)                                                                   \
  do {                                                                                      \
    ncclResult_t r__ = (expr);                                                              \
    if (r__ != 0) IPS_FAIL(IPS_ERR_NCCL, "%s failed: %s", #expr, nccl().GetErrorString(r__)); \
  } while (0)

__global__ void set_count_kernel(int64_t* dst, int64_t v) { *dst = v; }

}  // namespace ips

using namespace ips;

extern "C" int ips_comm_unique_id_bytes(void) { return (int)sizeof(ncclUniqueId); }

extern "C" int ips_comm_unique_id(void* out, int bytes) {
  if (out == nullptr || bytes < (int)sizeof(ncclUniqueId))
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_comm_unique_id: need a %zu-byte buffer", sizeof(ncclUniqueId));
  if (!nccl().ok) IPS_FAIL(IPS_ERR_NCCL, "ips_comm_unique_id: libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  IPS_NCCL_OK(nccl().GetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return IPS_OK;
}

extern "C" int ips_comm_create(void** out, const void* unique_id, int bytes, int rank, int world) {
  if (out == nullptr || unique_id == nullptr || bytes < (int)sizeof(ncclUniqueId))
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_comm_create: bad arguments");
  if (world < 1 || rank < 0 || rank >= world) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_comm_create: bad rank %d of %d", rank, world);
  if (!nccl().ok) IPS_FAIL(IPS_ERR_NCCL, "ips_comm_create: libnccl.so.2 could not be loaded");
  Comm* c = new Comm();
  c->rank = rank;
  c->world = world;
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclResult_t r = nccl().CommInitRank(&c->nccl, world, id, rank);
  if (r != 0) {
    delete c;
    IPS_FAIL(IPS_ERR_NCCL, "ncclCommInitRank failed: %s", nccl().GetErrorString(r));
  }
  cudaError_t e = cudaMalloc(&c->d_counts, (size_t)world * sizeof(int64_t));
  if (e != cudaSuccess) {
    nccl().CommDestroy(c->nccl);
    delete c;
    IPS_FAIL(IPS_ERR_NOMEM, "ips_comm_create: %s", cudaGetErrorString(e));
  }
  *out = c;
  return IPS_OK;
}

extern "C" int ips_comm_destroy(void* comm) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (c == nullptr) return IPS_OK;
  cudaFree(c->d_counts);
  if (c->nccl) nccl().CommDestroy(c->nccl);
  delete c;
  return IPS_OK;
}

// all_rows [world][cap_per_rank][row_bytes]: rank r's rows land in block r (first counts[r]
// rows valid).  counts_dev [world] int64 on the device receives every rank's row count.
extern "C" int ips_allgather_rows(void* comm, const void* local_rows, int64_t n_local, int row_bytes,
                                  void* all_rows, int64_t* counts_dev, int64_t cap_per_rank,
                                  ips_stream_t stream) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (c == nullptr || all_rows == nullptr || counts_dev == nullptr)
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_allgather_rows: NULL argument");
  if (n_local < 0 || n_local > cap_per_rank || row_bytes <= 0)
    IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_allgather_rows: n_local=%lld cap=%lld row_bytes=%d", (long long)n_local,
             (long long)cap_per_rank, row_bytes);
  if (n_local > 0 && local_rows == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_allgather_rows: NULL rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t block = (size_t)cap_per_rank * row_bytes;
  char* mine = reinterpret_cast<char*>(all_rows) + (size_t)c->rank * block;
  // stage the local rows in their own block (in place all-gather), then one ncclAllGather
  if (n_local > 0 && mine != local_rows)
    IPS_CUDA_OK(cudaMemcpyAsync(mine, local_rows, (size_t)n_local * row_bytes, cudaMemcpyDeviceToDevice, st));
  set_count_kernel<<<1, 1, 0, st>>>(c->d_counts + c->rank, n_local);
  IPS_LAUNCH_OK("set_count_kernel");
  // counts and rows leave in one NCCL group (one launch)
  const bool grouped = nccl().GroupStart != nullptr && nccl().GroupEnd != nullptr;
  if (grouped) IPS_NCCL_OK(nccl().GroupStart());
  IPS_NCCL_OK(nccl().AllGather(c->d_counts + c->rank, c->d_counts, sizeof(int64_t), ncclInt8, c->nccl, st));
  IPS_NCCL_OK(nccl().AllGather(mine, all_rows, block, ncclInt8, c->nccl, st));
  if (grouped) IPS_NCCL_OK(nccl().GroupEnd());
  IPS_CUDA_OK(cudaMemcpyAsync(counts_dev, c->d_counts, (size_t)c->world * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  return IPS_OK;
}

// The form the plate pipeline uses: table [world][block_rows][row_bytes], rank r's block already
// in place (ips_pack_rows_block: header row with the row count, then the rows).  ONE ncclAllGather
// of fixed size; the counts travel in the headers, so the launching thread never waits for the
// device.  A caller that knows its row count may pass block_rows = 1 + that count (all ranks the
// same value) instead of the padded capacity.
extern "C" int ips_allgather_blocks(void* comm, void* table, int64_t block_rows, int row_bytes, ips_stream_t stream) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (c == nullptr || table == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_allgather_blocks: NULL argument");
  if (block_rows < 1 || row_bytes < 8) IPS_FAIL(IPS_ERR_BAD_SHAPE, "ips_allgather_blocks: block_rows=%lld row_bytes=%d",
                                                (long long)block_rows, row_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t block = (size_t)block_rows * row_bytes;
  char* mine = reinterpret_cast<char*>(table) + (size_t)c->rank * block;
  IPS_NCCL_OK(nccl().AllGather(mine, table, block, ncclInt8, c->nccl, st));
  return IPS_OK;
}

extern "C" int ips_comm_rank(void* comm, int* rank, int* world) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (c == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_comm_rank: NULL communicator");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return IPS_OK;
}

// ---- peer-memory form of the same gather -------------------------------------------------------
// On an NVSwitch box every GPU can write every peer's memory at full NVLink rate, and a bulk
// all-gather is nothing but "every rank stores its block into every peer's table".  Done with the
// copy engines (cudaMemcpyAsync between the local table and the peers' tables, mapped through CUDA
// IPC) it needs no SM at all, whereas an ncclAllGather kernel holds 16-24 CTAs for as long as the
// slowest rank takes to arrive -- measured at N = 8: the fused field kernel runs 14 % slower while
// NCCL chunks overlap it (profiles/README.md).  The blocks are pushed as the plate progresses; the
// ONE NCCL all-gather that remains at the end of the plate (ips_allgather_blocks on the per-rank
// totals) is also the barrier: its kernel starts after this rank's pushes on the same stream and
// completes only when every rank's has started.
namespace ips {
struct PeerTable {
  int rank = 0, world = 1;
  void* own = nullptr;
  void* peer[64] = {nullptr};
  // one copy stream per peer: a single stream of peer copies runs on one copy engine (measured ~50 GB/s
  // at N = 8, slower than the rows are produced); one stream per peer keeps several engines and links busy
  cudaStream_t lane[64] = {nullptr};
  cudaEvent_t ready = nullptr, done[64] = {nullptr};
};
}  // namespace ips

extern "C" int ips_peer_table_close(void* table);
extern "C" int ips_ipc_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int ips_ipc_export(void* dev_ptr, void* handle_out, int bytes) {
  if (dev_ptr == nullptr || handle_out == nullptr || bytes < (int)sizeof(cudaIpcMemHandle_t))
    IPS_FAIL(IPS_ERR_BAD_ARG, "ips_ipc_export: need a device pointer and a %zu-byte buffer", sizeof(cudaIpcMemHandle_t));
  cudaIpcMemHandle_t h;
  IPS_CUDA_OK(cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(handle_out, &h, sizeof(h));
  return IPS_OK;
}

// handles [world][ips_ipc_handle_bytes()]: every rank's exported table (the allocation's base address);
// own_table: this rank's table.  All tables have the same size and layout.
extern "C" int ips_peer_table_open(void** out, const void* handles, int rank, int world, void* own_table) {
  if (out == nullptr || handles == nullptr || own_table == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_peer_table_open: NULL argument");
  if (world < 1 || world > 64 || rank < 0 || rank >= world) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_peer_table_open: bad rank %d of %d", rank, world);
  PeerTable* t = new PeerTable();
  t->rank = rank; t->world = world; t->own = own_table;
  const char* hb = reinterpret_cast<const char*>(handles);
  for (int p = 0; p < world; ++p) {
    if (p == rank) { t->peer[p] = own_table; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, hb + (size_t)p * sizeof(h), sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(&t->peer[p], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < p; ++q)
        if (q != rank && t->peer[q]) cudaIpcCloseMemHandle(t->peer[q]);
      delete t;
      (void)cudaGetLastError();
      IPS_FAIL(IPS_ERR_CUDA, "ips_peer_table_open: cudaIpcOpenMemHandle(rank %d) failed: %s", p, cudaGetErrorString(e));
    }
  }
  bool ok = cudaEventCreateWithFlags(&t->ready, cudaEventDisableTiming) == cudaSuccess;
  for (int p = 0; p < world && ok; ++p) {
    if (p == rank) continue;
    ok = cudaStreamCreateWithFlags(&t->lane[p], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&t->done[p], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    (void)cudaGetLastError();
    ips_peer_table_close(t);
    IPS_FAIL(IPS_ERR_CUDA, "ips_peer_table_open: could not create the copy streams");
  }
  *out = t;
  return IPS_OK;
}

extern "C" int ips_peer_table_close(void* table) {
  PeerTable* t = reinterpret_cast<PeerTable*>(table);
  if (t == nullptr) return IPS_OK;
  for (int p = 0; p < t->world; ++p) {
    if (p == t->rank) continue;
    if (t->lane[p]) { cudaStreamSynchronize(t->lane[p]); cudaStreamDestroy(t->lane[p]); }
    if (t->done[p]) cudaEventDestroy(t->done[p]);
    if (t->peer[p]) cudaIpcCloseMemHandle(t->peer[p]);
  }
  if (t->ready) cudaEventDestroy(t->ready);
  delete t;
  return IPS_OK;
}

// Store bytes [offset, offset + bytes) of this rank's table into the same place of every peer's
// table: `world - 1` copy-engine transfers, no kernel.  In `stream`'s order: the copies start when the
// work queued on `stream` so far is done and `stream` continues when they have landed; they run
// concurrently on one internal stream per peer.
extern "C" int ips_peer_push(void* table, size_t offset, size_t bytes, ips_stream_t stream) {
  PeerTable* t = reinterpret_cast<PeerTable*>(table);
  if (t == nullptr) IPS_FAIL(IPS_ERR_BAD_ARG, "ips_peer_push: NULL table");
  if (bytes == 0 || t->world == 1) return IPS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const char* src = reinterpret_cast<const char*>(t->own) + offset;
  IPS_CUDA_OK(cudaEventRecord(t->ready, st));
  for (int k = 1; k < t->world; ++k) {
    const int p = (t->rank + k) % t->world;        // every rank starts with a different peer
    IPS_CUDA_OK(cudaStreamWaitEvent(t->lane[p], t->ready, 0));
    IPS_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(t->peer[p]) + offset, src, bytes, cudaMemcpyDeviceToDevice, t->lane[p]));
    IPS_CUDA_OK(cudaEventRecord(t->done[p], t->lane[p]));
    IPS_CUDA_OK(cudaStreamWaitEvent(st, t->done[p], 0));
  }
  return IPS_OK;
}
