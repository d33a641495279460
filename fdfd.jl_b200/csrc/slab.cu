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
//
// Agglomerated coarse levels.  Below level `ka` (the first level whose GLOBAL grid has <= 2^18 points) a level is a few
// hundred KB and every kernel on it is latency-bound, so cutting it into slabs only adds exchanges (measured at 4096^2 on
// 2 GPUs: 134 exchanges per iteration, 88 of them on levels 3-6, ~20 us each).  Instead the slabs keep levels 0..ka only
// (halo H = 2^ka rows); on every visit of level ka the coarse right-hand side is all-gathered (slabs are equal contiguous
// row ranges, so rank order is row order), every rank runs the rest of the cycle redundantly on the whole coarse grid with
// the ordinary single-GPU multigrid (levels ka..L-1 of the global hierarchy, built from the all-gathered level-ka eps_r),
// and copies its window of the result back.  One collective per visit replaces 11 exchanges.
#include "comm.cuh"
#include "krylov.cuh"
#include <chrono>
#include <cmath>

namespace {

// out (nx x nrows) <- rows (off + j) mod ny of in (nx x ny): a slab's window (owned rows + halo) of a global coarse array
__global__ void k_window(const c64* __restrict__ in, c64* __restrict__ out, int64_t nx, int64_t ny, int64_t off, int64_t nrows,
                         const int* __restrict__ done) {
  if (done && *done) return;
  const int64_t n = nx * nrows;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / nx, ix = i - j * nx;
    const int64_t gj = ((off + j) % ny + ny) % ny;
    out[i] = in[ix + nx * gj];
  }
}

__global__ void k_scale_src_rows(int64_t n, c128 k, const c128* __restrict__ src, c128* __restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = k * src[i];
}

double wall_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct SlabSolver {
  fdfd_ctx* ctx = nullptr;
  fdfd_comm* comm = nullptr;
  fdfd_solve_opts_t o{};
  FineOp op;
  Multigrid<float> mg;        // levels 0..ka (or the whole hierarchy when ka < 0) on the slab's rows
  Multigrid<float> mgc;       // levels ka..L-1 on the WHOLE coarse grid, identical on every rank
  FineOp opg;                 // host-side description of the global operator (for mgc)
  int ka = -1;                // agglomeration level, -1: none
  KrylovWork w;
  int64_t Nx = 0, nyl = 0, y0 = 0, H = 0, nloc = 0;

  // refresh the halo rows of a level-l array (rows of nx_l elements of `elem` bytes)
  int halo(int l, void* buf, size_t elem) {
    const int64_t nx = l == 0 ? Nx : mg.lv[l].nx;
    const int64_t h = H >> l, ny = nyl >> l;
    char* p = (char*)buf;
    const size_t row = (size_t)nx * elem;
    return comm->exchange(ctx, p, p + (size_t)(h + ny) * row, p + (size_t)h * row, p + (size_t)ny * row, (size_t)h * row);
  }
  // ---- halo validity bookkeeping (communication avoidance).  The halo is H >> l rows deep but a stencil-type kernel
  // consumes only one or two rows of it, so a refresh is not needed after every kernel: vu[l] / vf[l] count how many halo
  // rows of the level-l iterate / right-hand side are still exact, every kernel shrinks them by its stencil reach, and a
  // refresh happens only when fewer than keep(l) rows are left (keep > the reach, so that the cut PML y-lines of
  // neighbouring slabs still overlap by a few exact rows).  Pure host arithmetic, identical on every rank.
  std::vector<int64_t> vu, vf;
  bool lazy = true;
  int64_t keep(int l) const { return lazy ? std::min<int64_t>(H >> l, std::max<int64_t>(2, 16 >> l)) : (H >> l); }
  int settle_u(int l) { if (vu[l] < keep(l)) { FDFD_TRY(halo(l, mg.lv[l].u.p, sizeof(c64))); vu[l] = H >> l; } return FDFD_OK; }
  int settle_f(int l) { if (vf[l] < keep(l)) { FDFD_TRY(halo(l, mg.lv[l].f.p, sizeof(c64))); vf[l] = H >> l; } return FDFD_OK; }
  int smooth(int l, bool zero, bool prolong) {
    FDFD_TRY(mg.smooth(l, zero, prolong));
    int64_t d = zero ? vf[l] : std::min(vu[l], vf[l]) - 1;              // zero guess: u = w f / C is pointwise
    if (prolong) d = std::min(d, 2 * vu[l + 1] - 3);                      // u + P u_c is exact 2 vu_c - 2 rows deep, the sweep takes one
    vu[l] = std::max<int64_t>(d, 0);
    return settle_u(l);
  }
  // level ka: gather the right-hand side, finish the cycle on the whole coarse grid, take back this slab's window
  int coarse_visit(bool zero, int kind) {
    MGLevel<float>& L = mg.lv[ka];
    MGLevel<float>& G = mgc.lv[ka];
    const int64_t nx = L.nx, nyo = nyl >> ka, hk = H >> ka;   // owned rows / halo rows of the slab on level ka
    if (zero) FDFD_TRY(comm->allgather(ctx, L.f.p + hk * nx, G.f.p, (size_t)(nyo * nx) * sizeof(c64)));
    FDFD_TRY(mgc.cycle(ka, zero, kind));
    const int64_t n = nx * (nyo + 2 * hk);
    k_window<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(
        G.u.p, L.u.p, nx, G.ny, (y0 >> ka) - hk, nyo + 2 * hk, mg.done);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    vu[ka] = hk;
    return FDFD_OK;
  }
  // Multigrid<T>::cycle with the exchanges in between (kind: 0 = V, 1 = F, 2 = W truncated at wdepth)
  int cycle(int l, bool zero, int kind) {
    const MGParams& prm = mg.prm;
    if (l == ka) return coarse_visit(zero, kind);
    if (l == (int)mg.lv.size() - 1) {
      for (int s = 0; s < std::max(1, prm.coarse_sweeps); ++s) FDFD_TRY(smooth(l, zero && s == 0, false));
      return FDFD_OK;
    }
    for (int s = 0; s < std::max(1, prm.nu1); ++s) FDFD_TRY(smooth(l, zero && s == 0, false));
    FDFD_TRY(mg.restrict_residual(l));
    // coarse row J takes the residual of fine rows 2J-1..2J+1, i.e. the iterate of rows 2J-2..2J+2
    vf[l + 1] = std::max<int64_t>(0, std::min((vu[l] - 2) / 2, (vf[l] - 1) / 2));
    if (l + 1 != ka) FDFD_TRY(settle_f(l + 1));   // level ka gathers its owned rows instead (needs vu[l] >= 2: keep(l) >= 2)
    if (kind == 2 && l < prm.wdepth) { FDFD_TRY(cycle(l + 1, true, 2)); FDFD_TRY(cycle(l + 1, false, 2)); }
    else if (kind == 1) { FDFD_TRY(cycle(l + 1, true, 1)); FDFD_TRY(cycle(l + 1, false, 0)); }
    else FDFD_TRY(cycle(l + 1, true, kind == 2 ? 0 : kind));
    for (int s = 0; s < prm.nu2; ++s) FDFD_TRY(smooth(l, false, s == 0));  // the first post-sweep applies the correction
    return FDFD_OK;
  }
  int precond(bool hold, const void** out) {
    vu.assign(mg.lv.size(), 0); vf.assign(mg.lv.size(), 0);
    FDFD_TRY(halo(0, mg.rhs(), sizeof(c64)));   // the Krylov update kernels wrote the right-hand side with zero halo rows
    vf[0] = H;
    FDFD_TRY(cycle(0, true, mg.prm.cycle));
    // the fp64 stencil that consumes the result reads one halo row, and that row must be the owner's value bit for bit
    // (v = A ph has to be A times ONE global vector or the BiCGSTAB recurrence drifts): inside the cycle a halo copy may
    // differ slightly from its owner where a cut PML y-line touched it, here it may not
    if (vu[0] < H) { FDFD_TRY(halo(0, mg.lv[0].u.p, sizeof(c64))); vu[0] = H; }
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
      if (S->ka >= 0) for (size_t l = S->ka; l < S->mgc.lv.size(); ++l) { v.push_back(S->mgc.lv[l].u.p); v.push_back(S->mgc.lv[l].tmp.p); }
    };
    k.set_state = [S](const std::vector<void*>& v) {
      size_t i = 0;
      for (auto& L : S->mg.lv) { L.u.p = (c64*)v[i++]; L.tmp.p = (c64*)v[i++]; }
      S->mg.spare.p = (c64*)v[i++];
      if (S->ka >= 0) for (size_t l = S->ka; l < S->mgc.lv.size(); ++l) { S->mgc.lv[l].u.p = (c64*)v[i++]; S->mgc.lv[l].tmp.p = (c64*)v[i++]; }
    };
    return k;
  }
};

}  // namespace

// depth of the slab hierarchy (levels 0..L-1 on the slab's rows) and the agglomeration level ka (-1: none); also used by slab_multi.cu
void slab_depth(const fdfd_grid_t& g, double omega, const MGParams& prm, int64_t nyl, int* nlev_out, int* ka_out) {
  const std::vector<std::pair<int64_t, int64_t>> sizes = mg_level_sizes(g, omega, prm, 0);
  auto fits = [&](int L) { const int64_t h = (int64_t)1 << (L - 1); return nyl % h == 0 && g.Ny % h == 0 && h <= nyl; };
  int L = (int)sizes.size();
  int64_t agg_points = (int64_t)1 << 18;
  if (const char* e = getenv("FDFD_SLAB_AGG_POINTS")) agg_points = atoll(e);   // 0 disables the agglomeration (diagnostics)
  int ka = -1;
  if (agg_points > 0) {
    for (int l = 1; l < L; ++l) if (sizes[l].first * sizes[l].second <= agg_points) { ka = l; break; }
    // no level small enough (or only a deeper one fits the slab height): agglomerate at the deepest level the slabs allow
    if (ka < 0) ka = L - 1;
    while (ka >= 1 && !fits(ka + 1)) --ka;
    if (ka < 1) ka = -1;
  }
  if (ka >= 1) { *nlev_out = ka + 1; *ka_out = ka; return; }
  while (L > 1 && !fits(L)) --L;
  *nlev_out = L; *ka_out = -1;
}

namespace {

int solve_slab(fdfd_ctx* ctx, fdfd_comm* comm, const fdfd_grid_t* g, double omega, const fdfd_c128* eps_rows,
               const fdfd_c128* src_rows, const fdfd_solve_opts_t* opts, fdfd_c128* fields_rows, fdfd_info_t* info) {
  const double t0 = wall_ms();
  SlabSolver S;
  S.ctx = ctx; S.comm = comm;
  if (opts) S.o = *opts; else fdfd_default_opts(&S.o);
  ARG_CHECK(ctx, (S.o.solver == FDFD_SOLVER_BICGSTAB || S.o.solver == FDFD_SOLVER_AUTO) && S.o.precond == FDFD_PRECOND_MG && S.o.mg_precision == FDFD_MG_F32,
            "the slab solve runs BiCGSTAB + fp32 multigrid only");
  ARG_CHECK(ctx, S.o.mg_nu2 >= 1, "the slab solve needs mg_nu2 >= 1");
  ARG_CHECK(ctx, g->Ny % comm->nranks == 0, "Ny must be divisible by the number of slabs");
  int64_t y0 = 0, nyl = 0;
  fdfd_slab_rows(g, comm->nranks, comm->rank, &y0, &nyl);
  const MGParams prm = mg_params_from(S.o);
  int nlev = 1, ka = -1;
  slab_depth(*g, omega, prm, nyl, &nlev, &ka);
  S.ka = ka; S.y0 = y0;
  // halo: a multiple of 2^(nlev-1) (index alignment of the levels).  The cut PML y-lines of neighbouring slabs overlap by
  // the halo; measured: with 8 rows the iteration count of 4 slabs grows by 15-25 %, with 64 rows it equals one GPU's
  int64_t H = (int64_t)1 << (nlev - 1);
  { int64_t want = 64; if (const char* e = getenv("FDFD_SLAB_HALO")) want = atoll(e); while (H * 2 <= want && H * 2 <= nyl) H *= 2; }
  const int64_t Nx = g->Nx, nloc = nyl + 2 * H, Nloc = Nx * nloc;
  S.Nx = Nx; S.nyl = nyl; S.H = H; S.nloc = nloc;
  if (const char* e = getenv("FDFD_SLAB_LAZY")) S.lazy = atoi(e) != 0;   // 0: refresh after every kernel (diagnostics)
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
    FDFD_TRY(S.op.build_slab(ctx, *g, FDFD_ORDER_FB, omega, reinterpret_cast<const fdfd_c128*>(eps_loc.p), y0, nyl, nlev, H));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
  }
  FDFD_TRY(S.w.alloc(ctx, Nloc, apply_num_blocks(Nx, nloc), S.o.maxit, false));
  FDFD_TRY(S.mg.setup(ctx, S.op, prm));
  ARG_CHECK(ctx, S.mg.levels() == nlev, "internal: multigrid depth differs from the slab halo depth");
  if (ka >= 1) {
    // level-ka eps_r of the whole grid (owned rows of every slab, rank order = row order), then the global coarse hierarchy
    const MGLevel<float>& Lk = S.mg.lv[ka];
    const int64_t nyo = nyl >> ka;
    DevBuf<c128> eps_k;
    CUDA_TRY(ctx, eps_k.alloc((size_t)Lk.nx * nyo * comm->nranks));
    FDFD_TRY(comm->allgather(ctx, Lk.eps.p + (H >> ka) * Lk.nx, eps_k.p, (size_t)(Lk.nx * nyo) * sizeof(c128)));
    S.opg.g = *g; S.opg.pol = FDFD_TM; S.opg.ordering = FDFD_ORDER_FB; S.opg.omega = omega; S.opg.omega_pml = omega;
    host_coef_fine(*g, omega, FDFD_ORDER_FB, 1.0 / (kMu0 * g->L0), S.opg.hc);
    FDFD_TRY(S.mgc.setup(ctx, S.opg, prm, ka, eps_k.p));
    ARG_CHECK(ctx, S.mgc.lv[ka].nx == Lk.nx && S.mgc.lv[ka].ny == nyo * comm->nranks, "internal: global coarse level size mismatch");
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
  }
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
  info->mg_levels = ka >= 1 ? S.mgc.levels() : nlev;

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
  // like fdfd_solve_driven: an unconverged solve is an error of the call (the fields are still copied out as the best
  // approximation).  Every rank sees the same flag (the convergence scalars come out of the allreduce), so no rank is left behind.
  if (info->flag != FDFD_OK) { fdfd_set_error(ctx, "slab Krylov solver stopped with flag %d after %d iterations, relres %.3e", info->flag, info->iters, info->relres); return info->flag; }
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
