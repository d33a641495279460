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
  final_reduce<kVecThreads, 4>(reinterpret_cast<const double*>(partials), nb, res);
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
           const c128* __restrict__ v, TP* __restrict__ pf) {
  if (sc->done) return;
  const c128 beta = sc->beta, omega = sc->omega;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kVecThreads) {
    const c128 pn = r[i] + beta * (p[i] - omega * v[i]);
    p[i] = pn;
    if (pf) pf[i] = TP(pn);
  }
}

template <typename TP>
__global__ void __launch_bounds__(kVecThreads)
k_s_update(int64_t N, const KScal* __restrict__ sc, const c128* __restrict__ r, const c128* __restrict__ v,
           c128* __restrict__ s, TP* __restrict__ sf) {
  if (sc->done) return;
  const c128 alpha = sc->alpha;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kVecThreads) {
    const c128 sn = r[i] - alpha * v[i];
    s[i] = sn;
    if (sf) sf[i] = TP(sn);
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

__device__ __forceinline__ bool finite2(c128 a) { return isfinite(a.x) && isfinite(a.y); }

// alpha = rho / <rhat, v>
__global__ void k_scal_alpha(const c128* __restrict__ partials, int nb, KScal* sc) {
  if (sc->done) return;
  double res[2];
  final_reduce<kVecThreads, 2>(reinterpret_cast<const double*>(partials), nb, res);
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
  final_reduce<kVecThreads, 4>(reinterpret_cast<const double*>(partials), nb, res);
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
  final_reduce<kVecThreads, 4>(reinterpret_cast<const double*>(partials), nb, res);
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

fdfd_problem::~fdfd_problem() {
  delete mgf; delete mgd;
  if (h_scal) cudaFreeHost(h_scal);
}

// one preconditioner application: in-vector was already written (in TP precision) into the MG rhs by the update
// kernel; returns pointer to the result
template <typename T>
static int mg_apply(Multigrid<T>* mg, bool hold, const cplx<T>** out) {
  FDFD_TRY(mg->apply(out));
  if (hold) {  // park the result in `spare` so the next cycle does not overwrite it
    std::swap(mg->lv[0].u.p, mg->spare.p);
    *out = mg->spare.p;
  }
  return FDFD_OK;
}

template <typename TP>
static int bicgstab_loop(fdfd_problem* P, fdfd_info_t* info, TP* mg_rhs, int (*prec)(fdfd_problem*, bool, const TP**)) {
  fdfd_ctx* ctx = P->ctx;
  cudaStream_t st = ctx->stream;
  const int64_t N = P->op.g.Nx * P->op.g.Ny;
  const bool te = P->op.pol == FDFD_TE;
  const OpView<double> A = P->op.view();
  const int nvb = P->nvec_blocks;
  const int nab = apply_num_blocks(A.nx, A.ny);
  const fdfd_solve_opts_t& o = P->opts;
  const int check_every = std::max(1, o.check_every);
  const int hist_len = (int)P->hist.n;
  KScal* sc = P->scal.p;
  const bool mgp = o.precond == FDFD_PRECOND_MG;

  k_init<<<nvb, kVecThreads, 0, st>>>(N, P->b.p, P->x.p, P->r.p, P->rhat.p, P->p.p, P->v.p, P->partials.p); KLAUNCH(ctx);
  k_scal_init<<<1, kVecThreads, 0, st>>>(P->partials.p, nvb, sc, o.tol, 0); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());

  int restarts = 0, flag = FDFD_OK;
  int it_enq = 0;
  double true_rel = 0.0;
  while (true) {
    // ---- enqueue a chunk of iterations
    for (int c = 0; c < check_every && it_enq < o.maxit; ++c, ++it_enq) {
      const TP* ph = nullptr; const TP* sh = nullptr;
      k_p_update<TP><<<nvb, kVecThreads, 0, st>>>(N, sc, P->r.p, P->p.p, P->v.p, mgp ? mg_rhs : (TP*)nullptr); KLAUNCH(ctx);
      FDFD_TRY(prec(P, true, &ph));
      DotSpec d1; d1.ndot = 1; d1.d0 = P->rhat.p; d1.partials = P->partials.p; d1.done = &sc->done;
      FDFD_TRY(launch_apply(ctx, A, te, ph, sizeof(TP) == sizeof(c64), P->v.p, d1));
      k_scal_alpha<<<1, kVecThreads, 0, st>>>(P->partials.p, nab, sc); KLAUNCH(ctx);
      k_s_update<TP><<<nvb, kVecThreads, 0, st>>>(N, sc, P->r.p, P->v.p, P->s.p, mgp ? mg_rhs : (TP*)nullptr); KLAUNCH(ctx);
      FDFD_TRY(prec(P, false, &sh));
      DotSpec d2; d2.ndot = 2; d2.d0 = P->s.p; d2.partials = P->partials.p; d2.done = &sc->done;
      FDFD_TRY(launch_apply(ctx, A, te, sh, sizeof(TP) == sizeof(c64), P->t.p, d2));
      k_scal_omega<<<1, kVecThreads, 0, st>>>(P->partials.p, nab, sc); KLAUNCH(ctx);
      k_xr_update<TP><<<nvb, kVecThreads, 0, st>>>(N, sc, P->x.p, ph, sh, P->s.p, P->t.p, P->r.p, P->rhat.p, P->partials.p); KLAUNCH(ctx);
      k_scal_rho<<<1, kVecThreads, 0, st>>>(P->partials.p, nvb, sc, P->hist.p, hist_len); KLAUNCH(ctx);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(P->h_scal, sc, sizeof(KScal), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const KScal h = *P->h_scal;
    if (o.verbose) fprintf(stderr, "[fdfd_b200] it %d relres %.3e%s\n", h.iter, std::sqrt(h.rr / h.bnorm2), h.breakdown ? " (breakdown)" : "");
    const bool out_of_its = it_enq >= o.maxit;
    if (!h.done && !out_of_its) continue;
    if (h.bnorm2 == 0.0) { true_rel = 0.0; break; }  // b = 0 -> x = 0
    // ---- true residual with the fp64 operator
    DotSpec d0;
    FDFD_TRY(launch_apply(ctx, A, te, P->x.p, false, P->t.p, d0));
    k_true_resid<<<nvb, kVecThreads, 0, st>>>(N, P->b.p, P->t.p, P->partials.p); KLAUNCH(ctx);
    k_scal_init<<<1, kVecThreads, 0, st>>>(P->partials.p, nvb, sc, o.tol, 2); KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(P->h_scal, sc, sizeof(KScal), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    true_rel = std::sqrt(P->h_scal->rr / P->h_scal->bnorm2);
    if (std::isfinite(true_rel) && true_rel <= o.tol) { flag = FDFD_OK; break; }
    if (out_of_its) { flag = FDFD_ERR_NOCONV; break; }
    if (restarts >= 8) { flag = h.breakdown ? FDFD_ERR_BREAKDOWN : FDFD_ERR_NOCONV; break; }
    if (!std::isfinite(true_rel)) { flag = FDFD_ERR_BREAKDOWN; break; }
    // ---- restart from the current x (cures breakdown and recurrence drift)
    ++restarts;
    k_restart<<<nvb, kVecThreads, 0, st>>>(N, P->b.p, P->t.p, P->r.p, P->rhat.p, P->p.p, P->v.p, P->partials.p); KLAUNCH(ctx);
    k_scal_init<<<1, kVecThreads, 0, st>>>(P->partials.p, nvb, sc, o.tol, 1); KLAUNCH(ctx);
  }
  info->iters = P->h_scal->iter;
  info->relres = true_rel;
  info->flag = flag;
  info->restarts = restarts;
  return FDFD_OK;
}

static int prec_mg32(fdfd_problem* P, bool hold, const c64** out) { return mg_apply<float>(P->mgf, hold, out); }
static int prec_mg64(fdfd_problem* P, bool hold, const c128** out) { return mg_apply<double>(P->mgd, hold, out); }
static int prec_none(fdfd_problem* P, bool hold, const c128** out) { *out = hold ? P->p.p : P->s.p; return FDFD_OK; }
static int prec_jacobi(fdfd_problem* P, bool hold, const c128** out) {
  c128* dst = hold ? P->ph.p : P->sh.p;
  const c128* src = hold ? P->p.p : P->s.p;
  const int blocks = P->nvec_blocks;
  if (P->op.pol == FDFD_TE) k_jacobi<true><<<blocks, 256, 0, P->ctx->stream>>>(P->op.view(), src, dst, P->scal.p);
  else k_jacobi<false><<<blocks, 256, 0, P->ctx->stream>>>(P->op.view(), src, dst, P->scal.p);
  P->ctx->launches++;
  *out = dst;
  return FDFD_OK;
}

int problem_solve_bicgstab(fdfd_problem* P, fdfd_info_t* info) {
  fdfd_ctx* ctx = P->ctx;
  cudaEvent_t e0, e1;
  CUDA_TRY(ctx, cudaEventCreate(&e0)); CUDA_TRY(ctx, cudaEventCreate(&e1));
  const int64_t l0 = ctx->launches;
  CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
  int st;
  if (P->opts.precond == FDFD_PRECOND_MG) {
    if (P->mgf) st = bicgstab_loop<c64>(P, info, P->mgf->rhs(), prec_mg32);
    else st = bicgstab_loop<c128>(P, info, P->mgd->rhs(), prec_mg64);
  } else if (P->opts.precond == FDFD_PRECOND_JACOBI) {
    st = bicgstab_loop<c128>(P, info, nullptr, prec_jacobi);
  } else {
    st = bicgstab_loop<c128>(P, info, nullptr, prec_none);
  }
  cudaEventRecord(e1, ctx->stream);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  info->solve_ms = ms;
  info->launches = ctx->launches - l0;
  info->setup_ms = P->setup_ms;
  info->mg_levels = P->mgf ? P->mgf->levels() : (P->mgd ? P->mgd->levels() : 0);
  P->have_x = (st == FDFD_OK);
  return st;
}
