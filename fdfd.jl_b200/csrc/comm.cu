// comm.cu -- transports of the slab-sharded solve (see comm.cuh).
#include "comm.cuh"
#include <chrono>
#include <dlfcn.h>
#include <nccl.h>   // types only: the functions are bound with dlsym so the library loads without NCCL

struct fdfd_comm_group { CommGroup g; };

namespace {

// ---- NCCL bound at run time -------------------------------------------------------------------------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    // the soname resolves to the copy the host process already mapped (torch bundles one), else the system one
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.h) break; }
    if (!api.h) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define BIND(field, sym) do { *(void**)(&api.field) = dlsym(api.h, sym); if (!api.field) { api.err = std::string("libnccl lacks ") + sym; api.h = nullptr; return; } } while (0)
    BIND(GetUniqueId, "ncclGetUniqueId"); BIND(CommInitRank, "ncclCommInitRank"); BIND(CommDestroy, "ncclCommDestroy");
    BIND(AllReduce, "ncclAllReduce"); BIND(AllGather, "ncclAllGather"); BIND(Send, "ncclSend"); BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart"); BIND(GroupEnd, "ncclGroupEnd"); BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
  });
  return &api;
}

#define NCCL_TRY(ctx, expr)                                                                          \
  do {                                                                                               \
    ncclResult_t _r = (expr);                                                                        \
    if (_r != ncclSuccess) {                                                                         \
      fdfd_set_error(ctx, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nccl_api()->GetErrorString(_r)); \
      return FDFD_ERR_CUDA;                                                                          \
    }                                                                                                \
  } while (0)

}  // namespace

void CommGroup::barrier() {
  std::unique_lock<std::mutex> lk(mu);
  const uint64_t gen = generation;
  if (++waiting == nranks) { waiting = 0; ++generation; cv.notify_all(); return; }
  // a member that failed never arrives: give up after a generous timeout instead of hanging the process
  if (!cv.wait_for(lk, std::chrono::seconds(600), [&] { return generation != gen || failed; })) failed = true;
  if (failed) cv.notify_all();
}

int fdfd_comm::exchange(fdfd_ctx* ctx, void* lo_halo, void* hi_halo, const void* lo_src, const void* hi_src, size_t bytes) {
  ++n_exchange; bytes_sent += 2 * (int64_t)bytes;
  cudaStream_t st = ctx->stream;
  if (nranks == 1) {  // the ring closes on itself: periodic wrap
    CUDA_TRY(ctx, cudaMemcpyAsync(lo_halo, hi_src, bytes, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(hi_halo, lo_src, bytes, cudaMemcpyDeviceToDevice, st));
    return FDFD_OK;
  }
  const int prev = (rank + nranks - 1) % nranks, next = (rank + 1) % nranks;
  if (kind == FDFD_COMM_NCCL) {
    NcclApi* a = nccl_api();
    ncclComm_t c = (ncclComm_t)nccl;
    // order matters when prev == next (2 ranks): the peer's first send (its hi rows) must meet my first recv (lo halo)
    NCCL_TRY(ctx, a->GroupStart());
    NCCL_TRY(ctx, a->Send(hi_src, bytes, ncclChar, next, c, st));
    NCCL_TRY(ctx, a->Send(lo_src, bytes, ncclChar, prev, c, st));
    NCCL_TRY(ctx, a->Recv(lo_halo, bytes, ncclChar, prev, c, st));
    NCCL_TRY(ctx, a->Recv(hi_halo, bytes, ncclChar, next, c, st));
    NCCL_TRY(ctx, a->GroupEnd());
    return FDFD_OK;
  }
  // threads: publish my sources once they are complete, pull from the neighbours, hold them until everybody has pulled
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  grp->lo_src[rank] = lo_src; grp->hi_src[rank] = hi_src;
  grp->barrier();
  if (grp->failed) { fdfd_set_error(ctx, "slab exchange: another rank failed or timed out"); return FDFD_ERR_CUDA; }
  CUDA_TRY(ctx, cudaMemcpyAsync(lo_halo, grp->hi_src[prev], bytes, cudaMemcpyDefault, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(hi_halo, grp->lo_src[next], bytes, cudaMemcpyDefault, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  grp->barrier();
  if (grp->failed) { fdfd_set_error(ctx, "slab exchange: another rank failed or timed out"); return FDFD_ERR_CUDA; }
  return FDFD_OK;
}

int fdfd_comm::allgather(fdfd_ctx* ctx, const void* send, void* recv, size_t bytes) {
  ++n_allgather; bytes_sent += (int64_t)bytes * (nranks - 1);
  cudaStream_t st = ctx->stream;
  if (nranks == 1) { CUDA_TRY(ctx, cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, st)); return FDFD_OK; }
  if (kind == FDFD_COMM_NCCL) {
    NCCL_TRY(ctx, nccl_api()->AllGather(send, recv, bytes, ncclChar, (ncclComm_t)nccl, st));
    return FDFD_OK;
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  grp->lo_src[rank] = send;
  grp->barrier();
  if (grp->failed) { fdfd_set_error(ctx, "slab allgather: another rank failed or timed out"); return FDFD_ERR_CUDA; }
  for (int r = 0; r < nranks; ++r)
    CUDA_TRY(ctx, cudaMemcpyAsync((char*)recv + (size_t)r * bytes, grp->lo_src[r], bytes, cudaMemcpyDefault, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  grp->barrier();
  if (grp->failed) { fdfd_set_error(ctx, "slab allgather: another rank failed or timed out"); return FDFD_ERR_CUDA; }
  return FDFD_OK;
}

int fdfd_comm::allreduce_sum4(fdfd_ctx* ctx, double* dev4) {
  ++n_allreduce;
  if (nranks == 1) return FDFD_OK;
  cudaStream_t st = ctx->stream;
  if (kind == FDFD_COMM_NCCL) {
    NCCL_TRY(ctx, nccl_api()->AllReduce(dev4, dev4, 4, ncclDouble, ncclSum, (ncclComm_t)nccl, st));
    return FDFD_OK;
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(h_pinned, dev4, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  for (int k = 0; k < 4; ++k) grp->sums[(size_t)rank * 4 + k] = h_pinned[k];
  grp->barrier();
  if (grp->failed) { fdfd_set_error(ctx, "slab allreduce: another rank failed or timed out"); return FDFD_ERR_CUDA; }
  for (int k = 0; k < 4; ++k) {  // fixed rank order: bit-identical on every rank
    double s = 0.0;
    for (int r = 0; r < nranks; ++r) s += grp->sums[(size_t)r * 4 + k];
    h_pinned[k] = s;
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(dev4, h_pinned, 4 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  grp->barrier();
  if (grp->failed) { fdfd_set_error(ctx, "slab allreduce: another rank failed or timed out"); return FDFD_ERR_CUDA; }
  return FDFD_OK;
}

// ---- C ABI --------------------------------------------------------------------------------------------------
extern "C" int fdfd_comm_unique_id(void* id_bytes) {
  if (!id_bytes) { fdfd_set_error(nullptr, "fdfd_comm_unique_id: NULL argument"); return FDFD_ERR_ARG; }
  static_assert(sizeof(ncclUniqueId) == FDFD_COMM_ID_BYTES, "NCCL unique id size");
  NcclApi* a = nccl_api();
  if (!a->h) { fdfd_set_error(nullptr, "fdfd_comm_unique_id: %s", a->err.c_str()); return FDFD_ERR_CUDA; }
  ncclUniqueId id;
  NCCL_TRY(nullptr, a->GetUniqueId(&id));
  std::memcpy(id_bytes, &id, sizeof id);
  return FDFD_OK;
}

extern "C" int fdfd_comm_create_nccl(fdfd_ctx* ctx, int nranks, int rank, const void* id_bytes, fdfd_comm** out) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  ARG_CHECK(ctx, out != nullptr && id_bytes != nullptr, "NULL argument");
  ARG_CHECK(ctx, nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  *out = nullptr;
  NcclApi* a = nccl_api();
  if (!a->h) { fdfd_set_error(ctx, "fdfd_comm_create_nccl: %s", a->err.c_str()); return FDFD_ERR_CUDA; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  std::memcpy(&id, id_bytes, sizeof id);
  ncclComm_t c = nullptr;
  NCCL_TRY(ctx, a->CommInitRank(&c, nranks, id, rank));
  fdfd_comm* cm = new fdfd_comm();
  cm->kind = FDFD_COMM_NCCL; cm->nranks = nranks; cm->rank = rank; cm->nccl = (void*)c;
  *out = cm;
  return FDFD_OK;
}

extern "C" int fdfd_comm_group_create(int nranks, fdfd_comm_group** out) {
  if (!out || nranks < 1) { fdfd_set_error(nullptr, "fdfd_comm_group_create: bad arguments"); return FDFD_ERR_ARG; }
  fdfd_comm_group* g = new fdfd_comm_group();
  g->g.nranks = nranks;
  g->g.lo_src.assign(nranks, nullptr); g->g.hi_src.assign(nranks, nullptr); g->g.sums.assign((size_t)nranks * 4, 0.0);
  *out = g;
  return FDFD_OK;
}

extern "C" void fdfd_comm_group_destroy(fdfd_comm_group* grp) { delete grp; }

extern "C" int fdfd_comm_create_threads(fdfd_comm_group* grp, int rank, fdfd_comm** out) {
  if (!grp || !out || rank < 0 || rank >= grp->g.nranks) { fdfd_set_error(nullptr, "fdfd_comm_create_threads: bad arguments"); return FDFD_ERR_ARG; }
  fdfd_comm* cm = new fdfd_comm();
  cm->kind = FDFD_COMM_THREADS; cm->nranks = grp->g.nranks; cm->rank = rank; cm->grp = &grp->g;
  if (cudaMallocHost((void**)&cm->h_pinned, 4 * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    delete cm;
    fdfd_set_error(nullptr, "fdfd_comm_create_threads: cudaMallocHost failed (no CUDA device?)");
    return FDFD_ERR_CUDA;
  }
  *out = cm;
  return FDFD_OK;
}

extern "C" void fdfd_comm_destroy(fdfd_comm* comm) {
  if (!comm) return;
  if (comm->kind == FDFD_COMM_NCCL && comm->nccl && nccl_api()->h) nccl_api()->CommDestroy((ncclComm_t)comm->nccl);
  if (comm->h_pinned) cudaFreeHost(comm->h_pinned);
  delete comm;
}

extern "C" int fdfd_comm_stats(fdfd_comm* comm, int64_t* n_exchange, int64_t* n_allreduce, int64_t* bytes_sent) {
  if (!comm) return FDFD_ERR_ARG;
  if (n_exchange) *n_exchange = comm->n_exchange;
  if (n_allreduce) *n_allreduce = comm->n_allreduce;
  if (bytes_sent) *bytes_sent = comm->bytes_sent;
  return FDFD_OK;
}
