// comm.cuh -- the two exchange steps of the row-slab sharded solve ("one large grid"):
//   * ring halo exchange of whole rows between neighbouring y-slabs (the operators are periodic, grid.jl:147,150,
//     so the ring closes between the last and the first slab), and
//   * the sum of <= 4 doubles over the ranks (Krylov dot products and norms).
// Two transports behind one interface:
//   FDFD_COMM_NCCL     one process per GPU; ncclSend/ncclRecv + ncclAllReduce on the solver's stream (NVLink / NVSwitch),
//                      stream-ordered, so a whole BiCGSTAB iteration incl. its exchanges is captured in one CUDA graph.
//                      NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded by the host process).
//   FDFD_COMM_THREADS  several slabs inside ONE process (one host thread per slab, same or different GPUs):
//                      device-to-device copies between the slabs' buffers with host barriers.  This is the transport of
//                      the single-GPU parity tests of the slab path; it is not graph-capturable.
#pragma once
#include "common.cuh"
#include <condition_variable>
#include <mutex>

struct CommGroup {   // shared by the threads of a FDFD_COMM_THREADS communicator
  int nranks = 0;
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  uint64_t generation = 0;
  std::vector<const void*> lo_src, hi_src;   // published per rank for the current exchange
  std::vector<double> sums;                  // [rank][4]
  bool failed = false;                       // a member hit an error: everybody bails out of the barriers
  void barrier();
};

struct fdfd_comm {
  int kind = FDFD_COMM_THREADS;
  int nranks = 1, rank = 0;
  void* nccl = nullptr;        // ncclComm_t
  CommGroup* grp = nullptr;
  double* h_pinned = nullptr;  // 4 doubles, thread transport
  int64_t n_exchange = 0, n_allreduce = 0, n_allgather = 0, bytes_sent = 0;
  bool capturable() const { return kind == FDFD_COMM_NCCL || nranks == 1; }

  // my `lo_src` rows go to the previous slab's high halo, my `hi_src` rows to the next slab's low halo;
  // `lo_halo` receives the previous slab's hi_src, `hi_halo` the next slab's lo_src.  Stream-ordered on ctx->stream.
  int exchange(fdfd_ctx* ctx, void* lo_halo, void* hi_halo, const void* lo_src, const void* hi_src, size_t bytes);
  // recv (nranks * bytes) <- concatenation in rank order of every rank's `send` (bytes each): the agglomerated coarse
  // level of the slab multigrid (slabs are equal contiguous row ranges, so rank order IS global row order)
  int allgather(fdfd_ctx* ctx, const void* send, void* recv, size_t bytes);
  // in-place sum over the ranks of 4 doubles in device memory; every rank ends with bit-identical values
  int allreduce_sum4(fdfd_ctx* ctx, double* dev4);
};
