// slab.cu -- K11: solve(d::Device, TM) (src/solver/driven.jl:4-59) for ONE large grid split into row slabs over
// several GPUs (BASELINE config 5).  New work: the reference has no parallel path.
//
// Layout.  Slab r owns global rows [y0, y0+nyl) and stores every grid array as Nx x (nyl + 2H) with H halo rows on each
// side, H = 2^(levels-1).  Multigrid level l of a slab has (nyl >> l) + 2 (H >> l) rows, so local coarse row J sits on
// local fine row 2J exactly as on a whole grid and every single-GPU kernel (stencil, smoothers, transfers, line
// relaxation) runs unchanged on the local arrays; their periodic wrap only ever touches the outermost halo row, whose
// value is never used.  x is never cut: a warp still reads 512 contiguous bytes and the x-wrap is local.
//
// Exchanges (comm.cuh).  Halo rows of a level-l array are refreshed from the two neighbouring slabs after every
// stencil-type kernel that writes it: each smoothing sweep (iterate u_l), each restriction (coarse right-hand side
// f_{l+1}), and once per preconditioner application for the fine right-hand side.  The multigrid cycle therefore runs in
// lock step over the slabs and is the SAME cycle as on one GPU (same hierarchy, same transfers, same point smoother);
// the only difference is that the y-lines of the PML line relaxation are cut at the slab ends (each slab relaxes its
// lines over its rows + halo, an overlapping block version).  Krylov vectors carry zero halo rows, so the fused dot
// products need no masking; each dot product costs one allreduce of <= 4 doubles.  With NCCL everything is
// stream-ordered and one BiCGSTAB iteration, exchanges included, is replayed as a single CUDA graph.
#include "comm.cuh"
#include "krylov.cuh"
#include <chrono>
#include <cmath>

namespace {

__global__ void k_scale_src_rows(int64_t n, c128 k, const c128* __restrict__ src, c128* __restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = k * src[i];
}

double wall_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct SlabSolver {
  fdfd_ctx* ctx = nullptr;
  fdfd_comm* comm = nullptr;
  fdfd_solve_opts_t o{};
  FineOp op;
  Multigrid<float> mg;
  KrylovWork w;
  int64_t Nx = 0, nyl = 0, H = 0, nloc = 0;

  // refresh the halo rows of a level-l array (rows of nx_l elements of `elem` bytes)
  int halo(int l, void* buf, size_t elem) {
    const int64_t nx = l == 0 ? Nx : mg.lv[l].nx;
    const int64_t h = H >> l, ny = nyl >> l;
    char* p = (char*)buf;
    const size_t row = (size_t)nx * elem;
    return comm->exchange(ctx, p, p + (size_t)(h + ny) * row, p + (size_t)h * row, p + (size_t)ny * row, (size_t)h * row);
  }
  int smooth(int l, bool zero, bool prolong) {
    FDFD_TRY(mg.smooth(l, zero, prolong));
    return halo(l, mg.lv[l].u.p, sizeof(c64));
  }
  // Multigrid<T>::cycle with the exchanges in between (kind: 0 = V, 1 = F, 2 = W truncated at wdepth)
  int cycle(int l, bool zero, int kind) {
    const MGParams& prm = mg.prm;
    if (l == (int)mg.lv.size() - 1) {
      for (int s = 0; s < std::max(1, prm.coarse_sweeps); ++s) FDFD_TRY(smooth(l, zero && s == 0, false));
      return FDFD_OK;
    }
    for (int s = 0; s < std::max(1, prm.nu1); ++s) FDFD_TRY(smooth(l, zero && s == 0, false));
    FDFD_TRY(mg.restrict_residual(l));
    FDFD_TRY(halo(l + 1, mg.lv[l + 1].f.p, sizeof(c64)));
    if (kind == 2 && l < prm.wdepth) { FDFD_TRY(cycle(l + 1, true, 2)); FDFD_TRY(cycle(l + 1, false, 2)); }
    else if (kind == 1) { FDFD_TRY(cycle(l + 1, true, 1)); FDFD_TRY(cycle(l + 1, false, 0)); }
    else FDFD_TRY(cycle(l + 1, true, kind == 2 ? 0 : kind));
    for (int s = 0; s < prm.nu2; ++s) FDFD_TRY(smooth(l, false, s == 0));  // the first post-sweep applies the correction
    return FDFD_OK;
  }
  int precond(bool hold, const void** out) {
    FDFD_TRY(halo(0, mg.rhs(), sizeof(c64)));
    FDFD_TRY(cycle(0, true, mg.prm.cycle));
    const c64* res = mg.lv[0].u.p;
    if (hold) { std::swap(mg.lv[0].u.p, mg.spare.p); res = mg.spare.p; }
    *out = res;
    return FDFD_OK;
  }
  KrylovOps make_ops() {
    KrylovOps k;
    SlabSolver* S = this;
    k.nab = apply_num_blocks(Nx, nyl);   // the apply launches over the owned rows only
    k.prec_f32 = true; k.prec_rhs = mg.rhs(); k.fscale = mg.rhs_scale;
    k.apply = [S](const void* x, bool x_f32, c128* y, const DotSpec& ds) -> int {
      DotSpec d = ds; d.row_lo = S->H; d.row_hi = S->H + S->nyl;
      // preconditioned vectors come out of the cycle with fresh halos; fp64 Krylov vectors carry zero halos
      if (!x_f32) FDFD_TRY(S->halo(0, const_cast<void*>(x), sizeof(c128)));
      return launch_apply(S->ctx, S->op.view(), false, x, x_f32, y, d);
    };
    k.precond = [S](bool hold, const void** out) { return S->precond(hold, out); };
    k.allreduce = [S](double* dev4) { return S->comm->allreduce_sum4(S->ctx, dev4); };
    k.get_state = [S](std::vector<void*>& v) {
      v.clear();
      for (auto& L : S->mg.lv) { v.push_back(L.u.p); v.push_back(L.tmp.p); }
      v.push_back(S->mg.spare.p);
    };
    k.set_state = [S](const std::vector<void*>& v) {
      size_t i = 0;
      for (auto& L : S->mg.lv) { L.u.p = (c64*)v[i++]; L.tmp.p = (c64*)v[i++]; }
      S->mg.spare.p = (c64*)v[i++];
    };
    return k;
  }
};

int slab_depth(const fdfd_grid_t& g, double omega, const MGParams& prm, int64_t nyl) {
  int L = (int)mg_level_sizes(g, omega, prm, 0).size();
  while (L > 1) {
    const int64_t h = (int64_t)1 << (L - 1);
    if (nyl % h == 0 && g.Ny % h == 0 && h <= nyl) break;
    --L;
  }
  return L;
}

int solve_slab(fdfd_ctx* ctx, fdfd_comm* comm, const fdfd_grid_t* g, double omega, const fdfd_c128* eps_rows,
               const fdfd_c128* src_rows, const fdfd_solve_opts_t* opts, fdfd_c128* fields_rows, fdfd_info_t* info) {
  const double t0 = wall_ms();
  SlabSolver S;
  S.ctx = ctx; S.comm = comm;
  if (opts) S.o = *opts; else fdfd_default_opts(&S.o);
  ARG_CHECK(ctx, S.o.solver == FDFD_SOLVER_BICGSTAB && S.o.precond == FDFD_PRECOND_MG && S.o.mg_precision == FDFD_MG_F32,
            "the slab solve runs BiCGSTAB + fp32 multigrid only");
  ARG_CHECK(ctx, S.o.mg_nu2 >= 1, "the slab solve needs mg_nu2 >= 1");
  ARG_CHECK(ctx, g->Ny % comm->nranks == 0, "Ny must be divisible by the number of slabs");
  int64_t y0 = 0, nyl = 0;
  fdfd_slab_rows(g, comm->nranks, comm->rank, &y0, &nyl);
  const MGParams prm = mg_params_from(S.o);
  const int nlev = slab_depth(*g, omega, prm, nyl);
  const int64_t H = (int64_t)1 << (nlev - 1), Nx = g->Nx, nloc = nyl + 2 * H, Nloc = Nx * nloc;
  S.Nx = Nx; S.nyl = nyl; S.H = H; S.nloc = nloc;
  if (!comm->capturable()) S.o.use_graph = 0;   // host barriers cannot be captured
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;

  // eps_r on the local rows: owned rows from the caller, halo rows from the neighbours
  {
    DevBuf<c128> eps_loc;
    CUDA_TRY(ctx, eps_loc.alloc(Nloc));
    FDFD_TRY(fdfd_copy_in(ctx, eps_loc.p + H * Nx, eps_rows, (size_t)nyl * Nx * sizeof(c128)));
    // not S.halo(): the multigrid levels do not exist yet
    char* p = (char*)eps_loc.p; const size_t row = (size_t)Nx * sizeof(c128);
    FDFD_TRY(comm->exchange(ctx, p, p + (size_t)(H + nyl) * row, p + (size_t)H * row, p + (size_t)nyl * row, (size_t)H * row));
    FDFD_TRY(S.op.build_slab(ctx, *g, FDFD_ORDER_FB, omega, reinterpret_cast<const fdfd_c128*>(eps_loc.p), y0, nyl, nlev));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
  }
  FDFD_TRY(S.w.alloc(ctx, Nloc, apply_num_blocks(Nx, nloc), S.o.maxit, false));
  FDFD_TRY(S.mg.setup(ctx, S.op, prm));
  S.mg.done = &S.w.scal.p->done;
  ARG_CHECK(ctx, S.mg.levels() == nlev, "internal: multigrid depth differs from the slab halo depth");
  // b = 1im*ω*src on the owned rows (driven.jl:36), zero halo rows
  CUDA_TRY(ctx, cudaMemsetAsync(S.w.b.p, 0, (size_t)Nloc * sizeof(c128), st));
  FDFD_TRY(fdfd_copy_in(ctx, S.w.t.p, src_rows, (size_t)nyl * Nx * sizeof(c128)));
  k_scale_src_rows<<<S.w.nvec_blocks, 256, 0, st>>>(nyl * Nx, c128(0.0, omega), S.w.t.p, S.w.b.p + H * Nx); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaMemsetAsync(S.w.t.p, 0, (size_t)Nloc * sizeof(c128), st));
  // first use of the allreduce outside any graph capture (NCCL sets its channels up lazily)
  CUDA_TRY(ctx, cudaMemsetAsync(S.w.lsum.p, 0, 4 * sizeof(double), st));
  FDFD_TRY(comm->allreduce_sum4(ctx, S.w.lsum.p));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  const double setup_ms = wall_ms() - t0;

  fdfd_info_t local{};
  if (!info) info = &local;
  std::memset(info, 0, sizeof(*info));
  KrylovOps ops = S.make_ops();
  FDFD_TRY(krylov_bicgstab(ctx, S.w, ops, S.o, info));
  info->setup_ms = setup_ms;
  info->mg_levels = nlev;

  // Hx, Hy from backward differences of Ez (driven.jl:40-41) on the local rows, owned rows copied out
  FDFD_TRY(S.halo(0, S.w.x.p, sizeof(c128)));
  {
    DevBuf<c128> f3;
    CUDA_TRY(ctx, f3.alloc(3 * Nloc));
    FDFD_TRY(launch_recover(ctx, S.op, S.w.x.p, 0, std::complex<double>(omega, 0.0), 0, f3.p));
    for (int c = 0; c < 3; ++c)
      FDFD_TRY(fdfd_copy_out(ctx, fields_rows + (size_t)c * nyl * Nx, f3.p + (size_t)c * Nloc + H * Nx, (size_t)nyl * Nx * sizeof(c128)));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
  }
  info->total_ms = wall_ms() - t0;
  if (info->flag != FDFD_OK) fdfd_set_error(ctx, "slab Krylov solver stopped with flag %d after %d iterations, relres %.3e", info->flag, info->iters, info->relres);
  return FDFD_OK;
}

}  // namespace

extern "C" int fdfd_slab_rows(const fdfd_grid_t* g, int nranks, int rank, int64_t* y0, int64_t* nrows) {
  if (!g || nranks < 1 || rank < 0 || rank >= nranks || g->Ny % nranks != 0) {
    fdfd_set_error(nullptr, "fdfd_slab_rows: need 0 <= rank < nranks and Ny divisible by nranks");
    return FDFD_ERR_ARG;
  }
  const int64_t n = g->Ny / nranks;
  if (y0) *y0 = n * rank;
  if (nrows) *nrows = n;
  return FDFD_OK;
}

extern "C" int fdfd_solve_driven_slab(fdfd_ctx* ctx, fdfd_comm* comm, const fdfd_grid_t* g, double omega,
                                      const fdfd_c128* eps_r_rows, const fdfd_c128* src_rows, const fdfd_solve_opts_t* opts,
                                      fdfd_c128* fields_rows, fdfd_info_t* info) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  ARG_CHECK(ctx, comm != nullptr, "comm is NULL");
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, eps_r_rows && src_rows && fields_rows, "NULL argument");
  ARG_CHECK(ctx, omega > 0, "omega must be > 0");
  const int st = solve_slab(ctx, comm, g, omega, eps_r_rows, src_rows, opts, fields_rows, info);
  if (st != FDFD_OK && comm->grp) {  // release the other threads from their barriers
    std::lock_guard<std::mutex> lk(comm->grp->mu);
    comm->grp->failed = true;
    comm->grp->cv.notify_all();
  }
  return st;
}
