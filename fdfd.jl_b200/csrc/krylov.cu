// krylov.cu -- K7: GPU-resident right-preconditioned BiCGSTAB.  Replaces the sparse direct solve of
// dolinearsolve (src/solver/solver.jl:29-35).  Every scalar (rho, alpha, omega, beta, norms, flags) lives in
// HBM; dot products are fused into the stencil / update kernels as per-CTA partials (warp shuffles) and
// finished by one-CTA scalar kernels, so an iteration never round-trips to the host.  The host only polls a
// pinned mirror of the scalars every `check_every` iterations; kernels enqueued past convergence early-exit
// on the device-side `done` flag.
#include "krylov.cuh"
#include "reduce.cuh"
#include <chrono>
#include <cmath>

int apply_num_blocks(int64_t nx, int64_t ny);

namespace {

constexpr int kVecThreads = 256;
constexpr int kScalThreads = 1024;  // one-CTA scalar kernels: the apply kernel leaves up to 32768 per-CTA partials

__global__ void __launch_bounds__(kVecThreads)
k_init(int64_t N, const c128* __restrict__ b, c128* __restrict__ x, c128* __restrict__ r, c128* __restrict__ rhat,
       c128* __restrict__ p, c128* __restrict__ v, c128* __restrict__ partials) {
  double acc[4] = {0, 0, 0, 0};
  const c128 z(0.0, 0.0);
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kVecThreads) {
    const c128 bi = b[i];
    x[i] = z; r[i] = bi; rhat[i] = bi; p[i] = z; v[i] = z;
    acc[2] += norm2(bi);
  }
  block_reduce_store<kVecThreads, 4>(acc, reinterpret_cast<double*>(partials) + (size_t)blockIdx.x * 4);
}

// restart from the current x: r = b - t (t = A x), rhat = r, p = v = 0
__global__ void __launch_bounds__(kVecThreads)
k_restart(int64_t N, const c128* __restrict__ b, const c128* __restrict__ t, c128* __restrict__ r, c128* __restrict__ rhat,
          c128* __restrict__ p, c128* __restrict__ v, c128* __restrict__ partials) {
  double acc[4] = {0, 0, 0, 0};
  const c128 z(0.0, 0.0);
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kVecThreads) {
    const c128 ri = b[i] - t[i];
    r[i] = ri; rhat[i] = ri; p[i] = z; v[i] = z;
    acc[2] += norm2(ri);
  }
  block_reduce_store<kVecThreads, 4>(acc, reinterpret_cast<double*>(partials) + (size_t)blockIdx.x * 4);
}

// mode 0: fresh start (sets bnorm2); mode 1: restart (keeps bnorm2, tol2, iter); mode 2: true-residual probe (rr only)
__global__ void k_scal_init(const c128* __restrict__ partials, int nb, KScal* sc, double tol, int mode) {
  double res[4];
  final_reduce<kScalThreads, 4>(reinterpret_cast<const double*>(partials), nb, res);
  if (threadIdx.x == 0) {
    const double rr = res[2];
    if (mode == 2) { sc->rr = rr; return; }
    if (mode == 0) { sc->bnorm2 = rr; sc->tol2 = tol * tol; sc->iter = 0; }
    sc->rho = c128(rr, 0.0); sc->rho_old = c128(1.0, 0.0); sc->alpha = c128(1.0, 0.0); sc->omega = c128(1.0, 0.0);
    sc->beta = c128(0.0, 0.0); sc->rr = rr; sc->breakdown = 0;
    sc->done = (rr <= sc->tol2 * sc->bnorm2) ? 1 : 0;
  }
}

template <typename TP>
__global__ void __launch_bounds__(kVecThreads)
k_p_update(int64_t N, const KScal* __restrict__ sc, const c128* __restrict__ r, c128* __restrict__ p,
           const c128* __restrict__ v, TP* __restrict__ pf, double fscale) {
  if (sc->done) return;
  const c128 beta = sc->beta, omega = sc->omega;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kVecThreads) {
    const c128 pn = r[i] + beta * (p[i] - omega * v[i]);
    p[i] = pn;
    if (pf) pf[i] = TP(c128(fscale * pn.x, fscale * pn.y));
  }
}

template <typename TP>
__global__ void __launch_bounds__(kVecThreads)
k_s_update(int64_t N, const KScal* __restrict__ sc, const c128* __restrict__ r, const c128* __restrict__ v,
           c128* __restrict__ s, TP* __restrict__ sf, double fscale) {
  if (sc->done) return;
  const c128 alpha = sc->alpha;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kVecThreads) {
    const c128 sn = r[i] - alpha * v[i];
    s[i] = sn;
    if (sf) sf[i] = TP(c128(fscale * sn.x, fscale * sn.y));
  }
}

template <typename TP>
__global__ void __launch_bounds__(kVecThreads)
k_xr_update(int64_t N, const KScal* __restrict__ sc, c128* __restrict__ x, const TP* __restrict__ ph,
            const TP* __restrict__ sh, const c128* __restrict__ s, const c128* __restrict__ t, c128* __restrict__ r,
            const c128* __restrict__ rhat, c128* __restrict__ partials) {
  if (sc->done) return;
  const c128 alpha = sc->alpha, omega = sc->omega;
  double acc[4] = {0, 0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kVecThreads) {
    x[i] = x[i] + alpha * c128(ph[i]) + omega * c128(sh[i]);
    const c128 rn = s[i] - omega * t[i];
    r[i] = rn;
    const c128 q = cmulc(rhat[i], rn);
    acc[0] += q.x; acc[1] += q.y; acc[2] += norm2(rn);
  }
  block_reduce_store<kVecThreads, 4>(acc, reinterpret_cast<double*>(partials) + (size_t)blockIdx.x * 4);
}

// slab mode: local partials -> 4 doubles (then summed over the ranks by an allreduce) ; K doubles per partial entry
template <int K>
__global__ void k_reduce_partials(const c128* __restrict__ partials, int nb, double* __restrict__ lsum) {
  double res[K];
  final_reduce<kScalThreads, K>(reinterpret_cast<const double*>(partials), nb, res);
  if (threadIdx.x == 0) { for (int k = 0; k < 4; ++k) lsum[k] = k < K ? res[k] : 0.0; }
}

__device__ __forceinline__ bool finite2(c128 a) { return isfinite(a.x) && isfinite(a.y); }

// alpha = rho / <rhat, v>
__global__ void k_scal_alpha(const c128* __restrict__ partials, int nb, KScal* sc) {
  if (sc->done) return;
  double res[2];
  final_reduce<kScalThreads, 2>(reinterpret_cast<const double*>(partials), nb, res);
  if (threadIdx.x == 0) {
    const c128 rhv(res[0], res[1]);
    if (!finite2(rhv) || norm2(rhv) == 0.0) { sc->breakdown = 1; sc->done = 1; return; }
    sc->alpha = cdiv(sc->rho, rhv);
  }
}

// omega = <t, s> / <t, t>
__global__ void k_scal_omega(const c128* __restrict__ partials, int nb, KScal* sc) {
  if (sc->done) return;
  double res[4];
  final_reduce<kScalThreads, 4>(reinterpret_cast<const double*>(partials), nb, res);
  if (threadIdx.x == 0) {
    const double tt = res[2];
    if (!(tt > 0.0) || !isfinite(tt)) { sc->omega = c128(0.0, 0.0); }
    else sc->omega = c128(res[0] / tt, res[1] / tt);
  }
}

// rho_new = <rhat, r>, rr = <r, r>, beta = (rho_new/rho)(alpha/omega); convergence / breakdown flags
__global__ void k_scal_rho(const c128* __restrict__ partials, int nb, KScal* sc, double* __restrict__ hist, int hist_len) {
  if (sc->done) return;
  double res[4];
  final_reduce<kScalThreads, 4>(reinterpret_cast<const double*>(partials), nb, res);
  if (threadIdx.x == 0) {
    const c128 rho_new(res[0], res[1]);
    const double rr = res[2];
    sc->iter += 1;
    sc->rr = rr;
    if (sc->iter < hist_len) hist[sc->iter] = rr;
    if (rr <= sc->tol2 * sc->bnorm2) { sc->done = 1; return; }
    const c128 beta = cdiv(rho_new, sc->rho) * cdiv(sc->alpha, sc->omega);
    if (!isfinite(rr) || !finite2(beta) || norm2(sc->omega) == 0.0 || norm2(rho_new) == 0.0) { sc->breakdown = 1; sc->done = 1; return; }
    sc->rho_old = sc->rho; sc->rho = rho_new; sc->beta = beta;
  }
}

// true residual probe: partial ||b - t||^2
__global__ void __launch_bounds__(kVecThreads)
k_true_resid(int64_t N, const c128* __restrict__ b, const c128* __restrict__ t, c128* __restrict__ partials) {
  double acc[4] = {0, 0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kVecThreads)
    acc[2] += norm2(b[i] - t[i]);
  block_reduce_store<kVecThreads, 4>(acc, reinterpret_cast<double*>(partials) + (size_t)blockIdx.x * 4);
}

// Jacobi preconditioner (fp64): out = in / diag(A)
template <bool TE>
__global__ void k_jacobi(OpView<double> op, const c128* __restrict__ in, c128* __restrict__ out, const KScal* __restrict__ sc) {
  if (sc->done) return;
  const int64_t N = op.nx * op.ny;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ix = n % op.nx, iy = n / op.nx;
    const int64_t ixp = ix + 1 == op.nx ? 0 : ix + 1, iyp = iy + 1 == op.ny ? 0 : iy + 1;
    c128 W = op.cxm[ix], E = op.cxp[ix], S = op.cym[iy], Nn = op.cyp[iy], m;
    if (TE) {
      W = W * op.gx[n]; E = E * op.gx[ixp + op.nx * iy]; S = S * op.gy[n]; Nn = Nn * op.gy[ix + op.nx * iyp];
      m = op.mass_const;
    } else m = op.mass[n];
    out[n] = cdiv(in[n], m - W - E - S - Nn);
  }
}

}  // namespace

int vec_blocks_for(fdfd_ctx* ctx, int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 8); }

int KrylovWork::alloc(fdfd_ctx* ctx, int64_t n_, int nparts, int maxit, bool jacobi_bufs) {
  n = n_;
  nvec_blocks = vec_blocks_for(ctx, n);
  nparts = std::max(nparts, nvec_blocks);
#define WALLOC(buf, cnt) do { if ((buf).alloc(cnt) != cudaSuccess) { cudaGetLastError(); fdfd_set_error(ctx, "out of device memory allocating %zu elements", (size_t)(cnt)); return FDFD_ERR_ALLOC; } } while (0)
  WALLOC(b, n); WALLOC(x, n); WALLOC(r, n); WALLOC(rhat, n); WALLOC(p, n); WALLOC(v, n); WALLOC(s, n); WALLOC(t, n);
  if (jacobi_bufs) { WALLOC(ph, n); WALLOC(sh, n); }
  WALLOC(partials, (size_t)nparts * 2);
  WALLOC(scal, 1);
  WALLOC(lsum, 4);
  WALLOC(hist, (size_t)std::max(16, maxit + 2));
#undef WALLOC
  if (cudaMallocHost((void**)&h_scal, sizeof(KScal)) != cudaSuccess) { fdfd_set_error(ctx, "cudaMallocHost failed"); return FDFD_ERR_ALLOC; }
  std::memset(h_scal, 0, sizeof(KScal));   // readable (iter = 0) before the first BiCGSTAB solve and after solves by other methods
  return FDFD_OK;
}

KrylovWork::~KrylovWork() {
  for (auto& kv : graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (h_scal) cudaFreeHost(h_scal);
}

fdfd_problem::~fdfd_problem() { mlkrylov_free(ml); delete mgf; delete mgd; }

int jacobi_apply(fdfd_ctx* ctx, const FineOp& op, const c128* in, c128* out, const int* done, int blocks) {
  // KScal::done sits at a fixed offset; the kernel takes the KScal pointer
  const KScal* sc = reinterpret_cast<const KScal*>(reinterpret_cast<const char*>(done) - offsetof(KScal, done));
  if (op.pol == FDFD_TE) k_jacobi<true><<<blocks, 256, 0, ctx->stream>>>(op.view(), in, out, sc);
  else k_jacobi<false><<<blocks, 256, 0, ctx->stream>>>(op.view(), in, out, sc);
  KLAUNCH(ctx);
  return FDFD_OK;
}

template <typename TP>
static int bicgstab_loop(fdfd_ctx* ctx, KrylovWork& W, const KrylovOps& ops, const fdfd_solve_opts_t& o, fdfd_info_t* info) {
  cudaStream_t st = ctx->stream;
  const int64_t N = W.n;
  const int nvb = W.nvec_blocks;
  const int nab = ops.nab;
  const int check_every = std::max(1, o.check_every);
  const int hist_len = (int)W.hist.n;
  KScal* sc = W.scal.p;
  TP* prhs = (TP*)ops.prec_rhs;
  const double fscale = ops.fscale;
  const bool use_graph = o.use_graph && st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;

  // where the scalar kernels read their sums: the per-CTA partials, or (slab mode) the 4 doubles left by a local reduce
  // followed by the allreduce over the ranks
  const c128* sp = W.partials.p; int snb = 0;
  auto sums = [&](int nb, int K) -> int {
    sp = W.partials.p; snb = nb;
    if (!ops.allreduce) return FDFD_OK;
    if (K == 2) k_reduce_partials<2><<<1, kScalThreads, 0, st>>>(W.partials.p, nb, W.lsum.p);
    else k_reduce_partials<4><<<1, kScalThreads, 0, st>>>(W.partials.p, nb, W.lsum.p);
    KLAUNCH(ctx);
    FDFD_TRY(ops.allreduce(W.lsum.p));
    sp = reinterpret_cast<const c128*>(W.lsum.p); snb = 1;
    return FDFD_OK;
  };
  k_init<<<nvb, kVecThreads, 0, st>>>(N, W.b.p, W.x.p, W.r.p, W.rhat.p, W.p.p, W.v.p, W.partials.p); KLAUNCH(ctx);
  FDFD_TRY(sums(nvb, 4));
  k_scal_init<<<1, kScalThreads, 0, st>>>(sp, snb, sc, o.tol, 0); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());

  int restarts = 0, flag = FDFD_OK;
  int it_enq = 0;
  double true_rel = 0.0;
  // stagnation watch (fp32 preconditioner noise can stall BiCGSTAB near 1e-8): every 400 iterations the best residual so far is
  // compared with the best of 400 iterations earlier; less than 10 % gained = stalled -> true-residual check + restart from x.
  // (Round 1 asked for a factor 2 per 400 iterations: the 16384^2 slab solve converges at ~1.5x per 400 iterations, was restarted
  // eight times for it -- every restart throws the Krylov space away -- and gave up at 6.2e-9 after 19919 iterations.)
  double best_rr = 1e300, mark_rr = 1e300; int mark_it = 0;
  while (true) {
    // ---- enqueue a chunk of iterations.  One iteration is a fixed kernel sequence (~100-250 launches, most of
    // them latency-bound coarse-level kernels), so it is captured once per buffer-rotation state into a CUDA
    // graph and replayed; the rotation (u/tmp ping-pong, spare) has a period of at most 6 iterations.
    for (int c = 0; c < check_every && it_enq < o.maxit; ++c, ++it_enq) {
      auto one_iteration = [&]() -> int {
        const void* ph = nullptr; const void* sh = nullptr;
        k_p_update<TP><<<nvb, kVecThreads, 0, st>>>(N, sc, W.r.p, W.p.p, W.v.p, prhs, fscale); KLAUNCH(ctx);
        FDFD_TRY(ops.precond(true, &ph));
        DotSpec d1; d1.ndot = 1; d1.d0 = W.rhat.p; d1.partials = W.partials.p; d1.done = &sc->done;
        FDFD_TRY(ops.apply(ph, sizeof(TP) == sizeof(c64), W.v.p, d1));
        FDFD_TRY(sums(nab, 2));
        k_scal_alpha<<<1, kScalThreads, 0, st>>>(sp, snb, sc); KLAUNCH(ctx);
        k_s_update<TP><<<nvb, kVecThreads, 0, st>>>(N, sc, W.r.p, W.v.p, W.s.p, prhs, fscale); KLAUNCH(ctx);
        FDFD_TRY(ops.precond(false, &sh));
        DotSpec d2; d2.ndot = 2; d2.d0 = W.s.p; d2.partials = W.partials.p; d2.done = &sc->done;
        FDFD_TRY(ops.apply(sh, sizeof(TP) == sizeof(c64), W.t.p, d2));
        FDFD_TRY(sums(nab, 4));
        k_scal_omega<<<1, kScalThreads, 0, st>>>(sp, snb, sc); KLAUNCH(ctx);
        k_xr_update<TP><<<nvb, kVecThreads, 0, st>>>(N, sc, W.x.p, (const TP*)ph, (const TP*)sh, W.s.p, W.t.p, W.r.p, W.rhat.p, W.partials.p); KLAUNCH(ctx);
        FDFD_TRY(sums(nvb, 4));
        k_scal_rho<<<1, kScalThreads, 0, st>>>(sp, snb, sc, W.hist.p, hist_len); KLAUNCH(ctx);
        return FDFD_OK;
      };
      if (!use_graph) { FDFD_TRY(one_iteration()); continue; }
      std::vector<void*> key;
      if (ops.get_state) ops.get_state(key);
      auto itg = W.graphs.find(key);
      if (itg == W.graphs.end()) {
        const int64_t l0 = ctx->launches;
        CUDA_TRY(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        int rc = one_iteration();
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc != FDFD_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        CUDA_TRY(ctx, ce);
        IterGraph ig;
        ig.nlaunch = ctx->launches - l0;
        ctx->launches = l0;
        CUDA_TRY(ctx, cudaGraphInstantiate(&ig.exec, graph, 0));
        cudaGraphDestroy(graph);
        if (ops.get_state) ops.get_state(ig.post);
        itg = W.graphs.emplace(key, ig).first;
      } else if (ops.set_state) {
        ops.set_state(itg->second.post);
      }
      CUDA_TRY(ctx, cudaGraphLaunch(itg->second.exec, st));
      ctx->launches += itg->second.nlaunch;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(W.h_scal, sc, sizeof(KScal), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const KScal h = *W.h_scal;
    if (o.verbose) fprintf(stderr, "[fdfd_b200] it %d relres %.3e%s\n", h.iter, std::sqrt(h.rr / h.bnorm2), h.breakdown ? " (breakdown)" : "");
    const bool out_of_its = it_enq >= o.maxit;
    if (h.rr < best_rr) best_rr = h.rr;
    bool stalled = false;
    if (h.iter - mark_it >= 400) {
      stalled = !h.done && !out_of_its && restarts < 8 && !(best_rr < 0.81 * mark_rr);
      mark_rr = best_rr; mark_it = h.iter;
    }
    if (!h.done && !out_of_its && !stalled) continue;
    if (h.bnorm2 == 0.0) { true_rel = 0.0; break; }  // b = 0 -> x = 0
    // ---- true residual with the fp64 operator
    DotSpec d0;
    FDFD_TRY(ops.apply(W.x.p, false, W.t.p, d0));
    k_true_resid<<<nvb, kVecThreads, 0, st>>>(N, W.b.p, W.t.p, W.partials.p); KLAUNCH(ctx);
    FDFD_TRY(sums(nvb, 4));
    k_scal_init<<<1, kScalThreads, 0, st>>>(sp, snb, sc, o.tol, 2); KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(W.h_scal, sc, sizeof(KScal), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    true_rel = std::sqrt(W.h_scal->rr / W.h_scal->bnorm2);
    if (std::isfinite(true_rel) && true_rel <= o.tol) { flag = FDFD_OK; break; }
    if (out_of_its) { flag = FDFD_ERR_NOCONV; break; }
    best_rr = 1e300; mark_rr = 1e300; mark_it = h.iter;
    if (restarts >= 8) { flag = h.breakdown ? FDFD_ERR_BREAKDOWN : FDFD_ERR_NOCONV; break; }
    if (!std::isfinite(true_rel)) { flag = FDFD_ERR_BREAKDOWN; break; }
    // ---- restart from the current x (cures breakdown and recurrence drift)
    ++restarts;
    k_restart<<<nvb, kVecThreads, 0, st>>>(N, W.b.p, W.t.p, W.r.p, W.rhat.p, W.p.p, W.v.p, W.partials.p); KLAUNCH(ctx);
    FDFD_TRY(sums(nvb, 4));
    k_scal_init<<<1, kScalThreads, 0, st>>>(sp, snb, sc, o.tol, 1); KLAUNCH(ctx);
  }
  info->iters = W.h_scal->iter;
  info->relres = true_rel;
  info->flag = flag;
  info->restarts = restarts;
  return FDFD_OK;
}

int krylov_bicgstab(fdfd_ctx* ctx, KrylovWork& W, const KrylovOps& ops, const fdfd_solve_opts_t& o, fdfd_info_t* info) {
  cudaEvent_t e0, e1;
  CUDA_TRY(ctx, cudaEventCreate(&e0)); CUDA_TRY(ctx, cudaEventCreate(&e1));
  const int64_t l0 = ctx->launches;
  CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
  int st = ops.prec_f32 ? bicgstab_loop<c64>(ctx, W, ops, o, info) : bicgstab_loop<c128>(ctx, W, ops, o, info);
  cudaEventRecord(e1, ctx->stream);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  info->solve_ms = ms;
  info->launches = ctx->launches - l0;
  return st;
}

// ---- the single-operator problem (driven TM / TE) -------------------------------------------------------
template <typename T> static void mg_get_state(Multigrid<T>* mg, std::vector<void*>& v) {
  for (auto& L : mg->lv) { v.push_back(L.u.p); v.push_back(L.tmp.p); }
  v.push_back(mg->spare.p);
}
template <typename T> static void mg_set_state(Multigrid<T>* mg, const std::vector<void*>& v, size_t& k) {
  for (auto& L : mg->lv) { L.u.p = (cplx<T>*)v[k++]; L.tmp.p = (cplx<T>*)v[k++]; }
  mg->spare.p = (cplx<T>*)v[k++];
}
// one preconditioner application: the input was already written (scaled, in T precision) into the MG rhs
template <typename T> static int mg_apply_hold(Multigrid<T>* mg, bool hold, const void** out) {
  const cplx<T>* res = nullptr;
  FDFD_TRY(mg->apply(&res));
  if (hold) {  // park the result in `spare` so the next cycle does not overwrite it
    std::swap(mg->lv[0].u.p, mg->spare.p);
    res = mg->spare.p;
  }
  *out = res;
  return FDFD_OK;
}

KrylovOps fdfd_problem::make_ops() {
  KrylovOps k;
  fdfd_problem* P = this;
  const bool te = op.pol == FDFD_TE;
  k.nab = apply_num_blocks(op.g.Nx, op.g.Ny);
  k.apply = [P, te](const void* x, bool x_f32, c128* y, const DotSpec& ds) { return launch_apply(P->ctx, P->op.view(), te, x, x_f32, y, ds); };
  if (opts.precond == FDFD_PRECOND_MG && mgf) {
    k.prec_f32 = true; k.prec_rhs = mgf->rhs(); k.fscale = mgf->rhs_scale;
    k.precond = [P](bool hold, const void** out) { return mg_apply_hold<float>(P->mgf, hold, out); };
    k.get_state = [P](std::vector<void*>& v) { v.clear(); mg_get_state(P->mgf, v); };
    k.set_state = [P](const std::vector<void*>& v) { size_t i = 0; mg_set_state(P->mgf, v, i); };
  } else if (opts.precond == FDFD_PRECOND_MG && mgd) {
    k.prec_f32 = false; k.prec_rhs = mgd->rhs(); k.fscale = mgd->rhs_scale;
    k.precond = [P](bool hold, const void** out) { return mg_apply_hold<double>(P->mgd, hold, out); };
    k.get_state = [P](std::vector<void*>& v) { v.clear(); mg_get_state(P->mgd, v); };
    k.set_state = [P](const std::vector<void*>& v) { size_t i = 0; mg_set_state(P->mgd, v, i); };
  } else if (opts.precond == FDFD_PRECOND_JACOBI) {
    k.prec_f32 = false;
    k.precond = [P](bool hold, const void** out) {
      c128* dst = hold ? P->w.ph.p : P->w.sh.p;
      *out = dst;
      return jacobi_apply(P->ctx, P->op, hold ? P->w.p.p : P->w.s.p, dst, &P->w.scal.p->done, P->w.nvec_blocks);
    };
  } else {
    k.prec_f32 = false;
    k.precond = [P](bool hold, const void** out) { *out = hold ? P->w.p.p : P->w.s.p; return FDFD_OK; };
  }
  return k;
}
