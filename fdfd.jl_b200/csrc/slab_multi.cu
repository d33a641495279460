// slab_multi.cu -- the row-slab sharded forms of
//   * solve(d::ModulatedDevice)   (src/solver/modulation.jl:35-119): nf = 2 ns + 1 coupled sidebands, and
//   * eigenfrequency(d, TM, nev)  (src/solver/eigen.jl:69-96): shift-invert Arnoldi whose inner solves are slab solves,
// for ONE large grid split into y-slabs over several GPUs.  New work: the reference has no parallel path.
//
// Both are the driven slab solve of slab.cu with `nf` operators side by side (nf = 1 for the eigenfrequency inner solve):
// same layout (every grid array of a slab is Nx x (nyl + 2H), H halo rows per side, level l keeps H >> l), same two
// exchange steps (ring halo refresh of whole rows; one sum of <= 4 doubles per Krylov dot product), same agglomerated
// coarse levels, same halo-validity bookkeeping.  The sidebands' coupling is pointwise (modulation.jl:95-98), so all
// sidebands of a grid point live on the same GPU and the coupling needs no exchange; the sidebands' slab vectors are
// slices of one (nf * Nloc) Krylov vector and their multigrid cycles run in lock step, so the halo bookkeeping is shared.
// The Arnoldi basis is sharded like every other vector (owned rows + zero halo rows); its dot products are the same
// allreduce, the Hessenberg matrix lives on the host (identical on every rank: the allreduce is bit-reproducible).
//
// GPU tests: tests/test_gpu_slab_multi.py (1 / 2 / 4 slabs against the single-GPU solves and the oracle).
#include "arnoldi.cuh"
#include "comm.cuh"
#include "krylov.cuh"
#include "reduce.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <memory>

using cd = std::complex<double>;

void slab_depth(const fdfd_grid_t& g, double omega, const MGParams& prm, int64_t nyl, int* nlev_out, int* ka_out);

double which_key(int which, cd nu);

namespace {

constexpr int kT = 256;

// out (nx x nrows) <- rows (off + j) mod ny of in (nx x ny): a slab's window (owned rows + halo) of a global coarse array
__global__ void k_window_m(const c64* __restrict__ in, c64* __restrict__ out, int64_t nx, int64_t ny, int64_t off, int64_t nrows,
                           const int* __restrict__ done) {
  if (done && *done) return;
  const int64_t n = nx * nrows;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / nx, ix = i - j * nx;
    const int64_t gj = ((off + j) % ny + ny) % ny;
    out[i] = in[ix + nx * gj];
  }
}
// b = k * src on n contiguous elements
__global__ void k_scale_rows(int64_t n, c128 k, const c128* __restrict__ src, c128* __restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = k * src[i];
}
// b = eps .* v * s  (inner right-hand side of the TM shift-invert step, eigen.cu k_eig_rhs)
__global__ void k_eps_scale(int64_t n, double s, const c128* __restrict__ eps, const c128* __restrict__ v, c128* __restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const c128 q = eps[i] * v[i];
    b[i] = c128(q.x * s, q.y * s);
  }
}
// partial sums of <a, b> = sum conj(a) b into lsum-style [block][4] doubles (entries 2,3 zero)
__global__ void __launch_bounds__(kT) k_dotc4(int64_t N, const c128* __restrict__ a, const c128* __restrict__ b, double* __restrict__ partials) {
  double acc[4] = {0, 0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < N; i += (int64_t)gridDim.x * kT) {
    const c128 q = cmulc(a[i], b[i]);
    acc[0] += q.x; acc[1] += q.y;
  }
  block_reduce_store<kT, 4>(acc, partials + (size_t)blockIdx.x * 4);
}
__global__ void k_sum4(const double* __restrict__ partials, int nb, double* __restrict__ out4) {
  double res[4];
  final_reduce<kT, 4>(partials, nb, res);
  if (threadIdx.x < 4) out4[threadIdx.x] = res[threadIdx.x];
}
// w -= c * v ;  out (+)= c * v ; out = s * w
__global__ void k_axpy_m(int64_t N, c128 c, const c128* __restrict__ v, c128* __restrict__ w) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) w[i] -= c * v[i];
}
__global__ void k_axpy_acc(int64_t N, c128 c, const c128* __restrict__ v, c128* __restrict__ out, int first) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = first ? c * v[i] : out[i] + c * v[i];
}
__global__ void k_scale_real(int64_t N, double s, const c128* __restrict__ w, c128* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = c128(w[i].x * s, w[i].y * s);
}
// deterministic pseudo-random start vector keyed on the GLOBAL point index, so the Arnoldi run does not depend on the
// number of slabs (same generator as eigen.cu k_seed); rows [row0, row0 + nrows) of the global grid -> out (nx x nrows)
__global__ void k_seed_rows(int64_t nx, int64_t nrows, int64_t row0, c128* __restrict__ out, uint64_t seed) {
  const int64_t n = nx * nrows;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gi = i + nx * row0;
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(gi + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    out[i] = c128((double)(z & 0xFFFFFFFF) / 4294967296.0 - 0.5, (double)(z >> 32) / 4294967296.0 - 0.5);
  }
}

double wall_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// nf TM operators on one slab, solved together by one BiCGSTAB.  nf = 1: driven / shift-invert inner solve.
struct SlabMulti {
  fdfd_ctx* ctx = nullptr;
  fdfd_comm* comm = nullptr;
  fdfd_solve_opts_t o{};
  fdfd_grid_t gg{};
  int nf = 1;
  std::vector<double> omegan;
  std::vector<std::unique_ptr<FineOp>> ops, opgs;           // slab operators; host-side descriptions of the global ones (for mgcs)
  std::vector<std::unique_ptr<Multigrid<float>>> mgs, mgcs;  // levels 0..ka on the slab rows; levels ka..L-1 on the whole coarse grid
  DevBuf<c128> deps;             // Deps_r on the local rows (nf > 1)
  DevBuf<c64> F, U, Tm, S;       // contiguous level-0 buffers of all sidebands (rhs, u, tmp, spare)
  double eps0 = 0;
  int ka = -1, nlev = 1;
  KrylovWork w;
  int64_t Nx = 0, nyl = 0, y0 = 0, H = 0, nloc = 0, Nloc = 0;
  double setup_ms = 0;

  // ---- exchanges ---------------------------------------------------------------------------------------------------
  int halo_rows(void* buf, int64_t nx, int64_t h, int64_t ny, size_t elem) {
    char* p = (char*)buf;
    const size_t row = (size_t)nx * elem;
    return comm->exchange(ctx, p, p + (size_t)(h + ny) * row, p + (size_t)h * row, p + (size_t)ny * row, (size_t)h * row);
  }
  int halo(int l, void* buf, size_t elem) { return halo_rows(buf, l == 0 ? Nx : mgs[0]->lv[l].nx, H >> l, nyl >> l, elem); }

  // ---- halo validity bookkeeping, shared by the sidebands (their cycles run in lock step); see slab.cu -------------------
  std::vector<int64_t> vu, vf;
  bool lazy = true;
  int64_t keep(int l) const { return lazy ? std::min<int64_t>(H >> l, std::max<int64_t>(2, 16 >> l)) : (H >> l); }
  int settle_u(int l) {
    if (vu[l] < keep(l)) { for (auto& mg : mgs) FDFD_TRY(halo(l, mg->lv[l].u.p, sizeof(c64))); vu[l] = H >> l; }
    return FDFD_OK;
  }
  int settle_f(int l) {
    if (vf[l] < keep(l)) { for (auto& mg : mgs) FDFD_TRY(halo(l, mg->lv[l].f.p, sizeof(c64))); vf[l] = H >> l; }
    return FDFD_OK;
  }
  int smooth(int l, bool zero, bool prolong) {
    for (auto& mg : mgs) FDFD_TRY(mg->smooth(l, zero, prolong));
    int64_t d = zero ? vf[l] : std::min(vu[l], vf[l]) - 1;
    if (prolong) d = std::min(d, 2 * vu[l + 1] - 3);
    vu[l] = std::max<int64_t>(d, 0);
    return settle_u(l);
  }
  int coarse_visit(bool zero, int kind) {
    for (int j = 0; j < nf; ++j) {
      MGLevel<float>& L = mgs[j]->lv[ka];
      MGLevel<float>& G = mgcs[j]->lv[ka];
      const int64_t nx = L.nx, nyo = nyl >> ka, hk = H >> ka;
      if (zero) FDFD_TRY(comm->allgather(ctx, L.f.p + hk * nx, G.f.p, (size_t)(nyo * nx) * sizeof(c64)));
      FDFD_TRY(mgcs[j]->cycle(ka, zero, kind));
      const int64_t n = nx * (nyo + 2 * hk);
      k_window_m<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(
          G.u.p, L.u.p, nx, G.ny, (y0 >> ka) - hk, nyo + 2 * hk, mgs[j]->done);
      KLAUNCH(ctx);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    vu[ka] = H >> ka;
    return FDFD_OK;
  }
  int cycle(int l, bool zero, int kind) {
    const MGParams& prm = mgs[0]->prm;
    if (l == ka) return coarse_visit(zero, kind);
    if (l == nlev - 1) {
      for (int s = 0; s < std::max(1, prm.coarse_sweeps); ++s) FDFD_TRY(smooth(l, zero && s == 0, false));
      return FDFD_OK;
    }
    for (int s = 0; s < std::max(1, prm.nu1); ++s) FDFD_TRY(smooth(l, zero && s == 0, false));
    for (auto& mg : mgs) FDFD_TRY(mg->restrict_residual(l));
    vf[l + 1] = std::max<int64_t>(0, std::min((vu[l] - 2) / 2, (vf[l] - 1) / 2));
    if (l + 1 != ka) FDFD_TRY(settle_f(l + 1));
    if (kind == 2 && l < prm.wdepth) { FDFD_TRY(cycle(l + 1, true, 2)); FDFD_TRY(cycle(l + 1, false, 2)); }
    else if (kind == 1) { FDFD_TRY(cycle(l + 1, true, 1)); FDFD_TRY(cycle(l + 1, false, 0)); }
    else FDFD_TRY(cycle(l + 1, true, kind == 2 ? 0 : kind));
    for (int s = 0; s < prm.nu2; ++s) FDFD_TRY(smooth(l, false, s == 0));
    return FDFD_OK;
  }
  int precond(bool hold, const void** out) {
    vu.assign(nlev, 0); vf.assign(nlev, 0);
    for (auto& mg : mgs) FDFD_TRY(halo(0, mg->rhs(), sizeof(c64)));
    vf[0] = H;
    FDFD_TRY(cycle(0, true, mgs[0]->prm.cycle));
    if (vu[0] < H) { for (auto& mg : mgs) FDFD_TRY(halo(0, mg->lv[0].u.p, sizeof(c64))); vu[0] = H; }
    const c64* res = mgs[0]->lv[0].u.p;
    if (hold) {
      for (auto& mg : mgs) std::swap(mg->lv[0].u.p, mg->spare.p);
      res = mgs[0]->spare.p;
    }
    *out = res;   // the sidebands' buffers rotate in lock step, so their results stay contiguous
    return FDFD_OK;
  }
  int apply(const void* x, bool x_f32, c128* y, const DotSpec& ds) {
    const int nab1 = apply_num_blocks(Nx, nyl);
    const size_t esz = x_f32 ? sizeof(c64) : sizeof(c128);
    // preconditioned vectors come out of the cycle with fresh halos; fp64 Krylov vectors carry stale / zero halos
    if (!x_f32) for (int j = 0; j < nf; ++j) FDFD_TRY(halo(0, (char*)const_cast<void*>(x) + (size_t)j * Nloc * esz, sizeof(c128)));
    for (int j = 0; j < nf; ++j) {
      DotSpec d = ds; d.row_lo = H; d.row_hi = H + nyl;
      if (ds.ndot > 0) { d.partials = ds.partials + (size_t)j * nab1 * ds.ndot; d.d0 = ds.d0 + (size_t)j * Nloc; }
      Coupling c;
      if (nf > 1) {
        c.deps = deps.p;
        c.hw = 0.5 * omegan[j] * omegan[j] * eps0;   // 0.5*ωn[j]^2 * (ϵ₀ L₀)  (modulation.jl:95-98)
        c.xm1 = j > 0 ? (const char*)x + (size_t)(j - 1) * Nloc * esz : nullptr;
        c.xp1 = j + 1 < nf ? (const char*)x + (size_t)(j + 1) * Nloc * esz : nullptr;
      }
      FDFD_TRY(launch_apply(ctx, ops[j]->view(), false, (const char*)x + (size_t)j * Nloc * esz, x_f32, y + (size_t)j * Nloc, d,
                            nf > 1 ? &c : nullptr));
    }
    return FDFD_OK;
  }
  KrylovOps make_ops() {
    KrylovOps k;
    SlabMulti* S = this;
    k.nab = nf * apply_num_blocks(Nx, nyl);
    k.prec_f32 = true; k.prec_rhs = F.p; k.fscale = mgs[0]->rhs_scale;
    k.apply = [S](const void* x, bool x_f32, c128* y, const DotSpec& ds) { return S->apply(x, x_f32, y, ds); };
    k.precond = [S](bool hold, const void** out) { return S->precond(hold, out); };
    k.allreduce = [S](double* dev4) { return S->comm->allreduce_sum4(S->ctx, dev4); };
    k.get_state = [S](std::vector<void*>& v) {
      v.clear();
      for (int j = 0; j < S->nf; ++j) {
        for (auto& L : S->mgs[j]->lv) { v.push_back(L.u.p); v.push_back(L.tmp.p); }
        v.push_back(S->mgs[j]->spare.p);
        if (S->ka >= 0) for (size_t l = S->ka; l < S->mgcs[j]->lv.size(); ++l) { v.push_back(S->mgcs[j]->lv[l].u.p); v.push_back(S->mgcs[j]->lv[l].tmp.p); }
      }
    };
    k.set_state = [S](const std::vector<void*>& v) {
      size_t i = 0;
      for (int j = 0; j < S->nf; ++j) {
        for (auto& L : S->mgs[j]->lv) { L.u.p = (c64*)v[i++]; L.tmp.p = (c64*)v[i++]; }
        S->mgs[j]->spare.p = (c64*)v[i++];
        if (S->ka >= 0) for (size_t l = S->ka; l < S->mgcs[j]->lv.size(); ++l) { S->mgcs[j]->lv[l].u.p = (c64*)v[i++]; S->mgcs[j]->lv[l].tmp.p = (c64*)v[i++]; }
      }
    };
    return k;
  }

  // ---- setup: operators, hierarchies, Krylov vectors.  omegas[j] / omegas_pml[j]: mass-term and PML frequency of sideband j
  int setup(fdfd_ctx* ctx_, fdfd_comm* comm_, const fdfd_grid_t* g, int ordering, const std::vector<double>& omegas,
            const std::vector<double>& omegas_pml, const fdfd_c128* eps_rows, const fdfd_c128* deps_rows,
            const fdfd_solve_opts_t* opts) {
    const double t0 = wall_ms();
    ctx = ctx_; comm = comm_; gg = *g; nf = (int)omegas.size(); omegan = omegas;
    if (opts) o = *opts; else fdfd_default_opts(&o);
    ARG_CHECK(ctx, (o.solver == FDFD_SOLVER_BICGSTAB || o.solver == FDFD_SOLVER_AUTO) && o.precond == FDFD_PRECOND_MG && o.mg_precision == FDFD_MG_F32,
              "the slab solves run BiCGSTAB + fp32 multigrid only");
    ARG_CHECK(ctx, o.mg_nu2 >= 1, "the slab solves need mg_nu2 >= 1");
    ARG_CHECK(ctx, g->Ny % comm->nranks == 0, "Ny must be divisible by the number of slabs");
    FDFD_TRY(fdfd_slab_rows(g, comm->nranks, comm->rank, &y0, &nyl));
    const MGParams prm = mg_params_from(o);
    // one depth for all sidebands: the hierarchy is truncated where k0 h >= kh_stop, so take the HIGHEST frequency (it
    // stops first) and force that depth on the others -- the lock-step cycle needs identical level structures
    const double wmax = *std::max_element(omegas.begin(), omegas.end());
    slab_depth(*g, wmax, prm, nyl, &nlev, &ka);
    H = (int64_t)1 << (nlev - 1);
    { int64_t want = 64; if (const char* e = getenv("FDFD_SLAB_HALO")) want = atoll(e); while (H * 2 <= want && H * 2 <= nyl) H *= 2; }
    Nx = g->Nx; nloc = nyl + 2 * H; Nloc = Nx * nloc;
    eps0 = kEps0 * g->L0;
    if (const char* e = getenv("FDFD_SLAB_LAZY")) lazy = atoi(e) != 0;
    if (!comm->capturable()) o.use_graph = 0;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    // eps_r (and Deps_r) on the local rows: owned rows from the caller, halo rows from the neighbours
    DevBuf<c128> eps_loc;
    CUDA_TRY(ctx, eps_loc.alloc(Nloc));
    FDFD_TRY(fdfd_copy_in(ctx, eps_loc.p + H * Nx, eps_rows, (size_t)nyl * Nx * sizeof(c128)));
    FDFD_TRY(halo_rows(eps_loc.p, Nx, H, nyl, sizeof(c128)));
    if (nf > 1) {
      ARG_CHECK(ctx, deps_rows != nullptr, "Deps_r is NULL");
      CUDA_TRY(ctx, deps.alloc(Nloc));
      CUDA_TRY(ctx, cudaMemsetAsync(deps.p, 0, (size_t)Nloc * sizeof(c128), st));   // the coupling only reads owned rows
      FDFD_TRY(fdfd_copy_in(ctx, deps.p + H * Nx, deps_rows, (size_t)nyl * Nx * sizeof(c128)));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    FDFD_TRY(w.alloc(ctx, (int64_t)nf * Nloc, nf * apply_num_blocks(Nx, nloc), o.maxit, false));
    CUDA_TRY(ctx, F.alloc((size_t)nf * Nloc)); CUDA_TRY(ctx, U.alloc((size_t)nf * Nloc));
    CUDA_TRY(ctx, Tm.alloc((size_t)nf * Nloc)); CUDA_TRY(ctx, S.alloc((size_t)nf * Nloc));
    for (int j = 0; j < nf; ++j) {
      ARG_CHECK(ctx, omegas[j] > 0 && omegas_pml[j] > 0, "a sideband frequency is not positive");
      ops.emplace_back(new FineOp());
      FDFD_TRY(ops[j]->build_slab(ctx, *g, ordering, omegas[j], reinterpret_cast<const fdfd_c128*>(eps_loc.p), y0, nyl, nlev, H, omegas_pml[j]));
      mgs.emplace_back(new Multigrid<float>());
      FDFD_TRY(mgs[j]->setup(ctx, *ops[j], prm));
      ARG_CHECK(ctx, mgs[j]->levels() == nlev, "internal: multigrid depth differs from the slab halo depth");
      MGLevel<float>& L0 = mgs[j]->lv[0];
      L0.f.alias(F.p + (size_t)j * Nloc, Nloc); L0.u.alias(U.p + (size_t)j * Nloc, Nloc); L0.tmp.alias(Tm.p + (size_t)j * Nloc, Nloc);
      mgs[j]->spare.alias(S.p + (size_t)j * Nloc, Nloc);
      if (ka >= 1) {
        const MGLevel<float>& Lk = mgs[j]->lv[ka];
        const int64_t nyo = nyl >> ka;
        DevBuf<c128> eps_k;
        CUDA_TRY(ctx, eps_k.alloc((size_t)Lk.nx * nyo * comm->nranks));
        FDFD_TRY(comm->allgather(ctx, Lk.eps.p + (H >> ka) * Lk.nx, eps_k.p, (size_t)(Lk.nx * nyo) * sizeof(c128)));
        opgs.emplace_back(new FineOp());
        FineOp& G = *opgs[j];
        G.g = *g; G.pol = FDFD_TM; G.ordering = ordering; G.omega = omegas[j]; G.omega_pml = omegas_pml[j];
        host_coef_fine(*g, omegas_pml[j], ordering, 1.0 / (kMu0 * g->L0), G.hc);
        mgcs.emplace_back(new Multigrid<float>());
        FDFD_TRY(mgcs[j]->setup(ctx, G, prm, ka, eps_k.p));
        ARG_CHECK(ctx, mgcs[j]->lv[ka].nx == Lk.nx && mgcs[j]->lv[ka].ny == nyo * comm->nranks, "internal: global coarse level size mismatch");
        CUDA_TRY(ctx, cudaStreamSynchronize(st));   // eps_k is a local
      }
    }
    CUDA_TRY(ctx, cudaMemsetAsync(w.b.p, 0, (size_t)nf * Nloc * sizeof(c128), st));
    CUDA_TRY(ctx, cudaMemsetAsync(w.t.p, 0, (size_t)nf * Nloc * sizeof(c128), st));
    // first use of the allreduce outside any graph capture (NCCL sets its channels up lazily)
    CUDA_TRY(ctx, cudaMemsetAsync(w.lsum.p, 0, 4 * sizeof(double), st));
    FDFD_TRY(comm->allreduce_sum4(ctx, w.lsum.p));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    setup_ms = wall_ms() - t0;
    return FDFD_OK;
  }
  // with the agglomerated coarse part the level-ka eps of the slab hierarchy must exist: Multigrid::setup allocates
  // L.eps for l >= 1 only, and ka >= 1 by construction (slab_depth)

  // solve with the right-hand side already in w.b (owned rows, zero halo rows).  t must carry zero halo rows.
  int solve(fdfd_info_t* info) {
    std::memset(info, 0, sizeof(*info));
    CUDA_TRY(ctx, cudaMemsetAsync(w.t.p, 0, (size_t)nf * Nloc * sizeof(c128), ctx->stream));
    KrylovOps ops_ = make_ops();
    FDFD_TRY(krylov_bicgstab(ctx, w, ops_, o, info));
    info->setup_ms = setup_ms;
    info->mg_levels = ka >= 1 ? mgcs[0]->levels() : nlev;
    return FDFD_OK;
  }
  // owned rows of sideband j of a local vector
  c128* owned(c128* v, int j) const { return v + (size_t)j * Nloc + H * Nx; }
  // (Nx, nyl, 3) field rows of sideband j from the solution x (forward: modulation.jl:112-113 / eigen.jl:90-91)
  int fields_out(int j, c128* xj_local, int forward, cd omega_field, fdfd_c128* fields_rows, DevBuf<c128>& f3) {
    FDFD_TRY(halo(0, xj_local, sizeof(c128)));
    if (!f3.p) CUDA_TRY(ctx, f3.alloc(3 * Nloc));
    FDFD_TRY(launch_recover(ctx, *ops[j], xj_local, forward, omega_field, 0, f3.p));
    for (int c = 0; c < 3; ++c)
      FDFD_TRY(fdfd_copy_out(ctx, fields_rows + (size_t)c * nyl * Nx, f3.p + (size_t)c * Nloc + H * Nx, (size_t)nyl * Nx * sizeof(c128)));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FDFD_OK;
  }
};

void release_group(fdfd_comm* comm, int st) {
  if (st != FDFD_OK && comm && comm->grp) {  // release the other threads from their barriers
    std::lock_guard<std::mutex> lk(comm->grp->mu);
    comm->grp->failed = true;
    comm->grp->cv.notify_all();
  }
}

int solve_modulated_slab(fdfd_ctx* ctx, fdfd_comm* comm, const fdfd_grid_t* g, double omega, double Omega, int ns, int sharedpml,
                         const fdfd_c128* eps_rows, const fdfd_c128* deps_rows, const fdfd_c128* src_rows,
                         const fdfd_solve_opts_t* opts, fdfd_c128* fields_rows, fdfd_info_t* info) {
  const double t0 = wall_ms();
  const int nf = 2 * ns + 1;
  std::vector<double> wn(nf), wp(nf);
  for (int j = 0; j < nf; ++j) {
    wn[j] = omega + Omega * (double)(j - ns);   // ωn = ω .+ Ω*n  (modulation.jl:41,49)
    ARG_CHECK(ctx, wn[j] > 0, "a sideband frequency w + n*Omega is not positive");
    wp[j] = sharedpml ? omega : wn[j];          // modulation.jl:79 / :87-91
  }
  SlabMulti S;
  FDFD_TRY(S.setup(ctx, comm, g, FDFD_ORDER_BF, wn, wp, eps_rows, deps_rows, opts));
  cudaStream_t st = ctx->stream;
  // b: zeros(N*nf); centre block = 1im*ω*src  (modulation.jl:67-69), owned rows only
  FDFD_TRY(fdfd_copy_in(ctx, S.w.t.p, src_rows, (size_t)S.nyl * S.Nx * sizeof(c128)));
  k_scale_rows<<<S.w.nvec_blocks, 256, 0, st>>>(S.nyl * S.Nx, c128(0.0, omega), S.w.t.p, S.owned(S.w.b.p, ns)); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  fdfd_info_t local{};
  if (!info) info = &local;
  FDFD_TRY(S.solve(info));
  DevBuf<c128> f3;
  for (int j = 0; j < nf; ++j)
    FDFD_TRY(S.fields_out(j, S.w.x.p + (size_t)j * S.Nloc, 1, cd(wn[j], 0.0), fields_rows + (size_t)j * 3 * S.nyl * S.Nx, f3));
  info->total_ms = wall_ms() - t0;
  if (info->flag != FDFD_OK) {
    fdfd_set_error(ctx, "fdfd_solve_modulated_slab: Krylov solver stopped with flag %d after %d iterations, relres %.3e", info->flag, info->iters, info->relres);
    return info->flag;
  }
  return FDFD_OK;
}

// sum over the ranks of <a, b> on the local vectors (zero halo rows); host value identical on every rank
int slab_dot(SlabMulti& S, DevBuf<double>& parts, const c128* a, const c128* b, cd* out) {
  fdfd_ctx* ctx = S.ctx;
  cudaStream_t st = ctx->stream;
  const int nb = S.w.nvec_blocks;
  k_dotc4<<<nb, kT, 0, st>>>(S.Nloc, a, b, parts.p); KLAUNCH(ctx);
  k_sum4<<<1, kT, 0, st>>>(parts.p, nb, S.w.lsum.p); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  FDFD_TRY(S.comm->allreduce_sum4(ctx, S.w.lsum.p));
  double h[4];
  CUDA_TRY(ctx, cudaMemcpyAsync(h, S.w.lsum.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  *out = cd(h[0], h[1]);
  return FDFD_OK;
}

int eigenfrequency_slab(fdfd_ctx* ctx, fdfd_comm* comm, const fdfd_grid_t* g, double omega0, int nev, int which, int ncv,
                        const fdfd_c128* eps_rows, const fdfd_solve_opts_t* opts, fdfd_c128* omega_out,
                        fdfd_c128* fields_rows, fdfd_info_t* info) {
  const double t0 = wall_ms();
  fdfd_solve_opts_t o;
  if (opts) o = *opts; else fdfd_default_opts(&o);
  o.tol = std::min(o.tol, 1e-11);          // inner solves; Ritz residual |b^T y| <= 10 o.tol |nu|   (as eigen.cu)
  const double tol_eig = 10.0 * o.tol;
  if (ncv <= 0) ncv = std::max(20, 2 * nev + 1);
  const int64_t Nglob = g->Nx * g->Ny;
  ARG_CHECK(ctx, nev + 2 <= Nglob, "nev too large for the grid");
  ncv = (int)std::min<int64_t>(std::max(ncv, nev + 2), Nglob - 1);
  const int max_steps = std::max(300, 30 * nev) + ncv;
  const double eps0 = kEps0 * g->L0, mu0 = kMu0 * g->L0;
  const cd sigma(-omega0 * omega0 * mu0 * eps0, 0);   // eigen.jl:86

  SlabMulti S;
  FDFD_TRY(S.setup(ctx, comm, g, FDFD_ORDER_FB, {omega0}, {omega0}, eps_rows, nullptr, &o));
  cudaStream_t st = ctx->stream;
  const int nb = S.w.nvec_blocks;
  const int64_t Nl = S.Nloc, own = S.nyl * S.Nx;
  std::vector<DevBuf<c128>> V;   // Arnoldi basis, sharded like every vector (owned rows + zero halo rows); at most ncv + 1 vectors
  auto ensure = [&](int j) -> int {
    while ((int)V.size() <= j) {
      V.emplace_back();
      if (V.back().alloc(Nl) != cudaSuccess) { cudaGetLastError(); fdfd_set_error(ctx, "out of device memory for the Arnoldi basis"); return FDFD_ERR_ALLOC; }
    }
    return FDFD_OK;
  };
  DevBuf<c128> wv; DevBuf<double> parts;
  CUDA_TRY(ctx, wv.alloc(Nl)); CUDA_TRY(ctx, parts.alloc((size_t)nb * 4));
  int inner_its = 0;
  const int64_t launches0 = ctx->launches;
  double inner_ms = 0;

  // the projected matrix lives on the host of every rank; it is identical everywhere because the dot products come out of the
  // bit-reproducible allreduce, so every rank takes the same restart decisions (no extra synchronisation)
  ArnoldiOps ops;
  ops.op_apply = [&](int j) -> int {
    // w = OP v_j:  (L/mu0 + w0^2 eps0 eps_r) y = eps_r v / mu0   -- the driven TM operator at w0
    k_eps_scale<<<nb, 256, 0, st>>>(Nl, 1.0 / mu0, S.ops[0]->eps.p, V[j].p, S.w.b.p); KLAUNCH(ctx);   // v has zero halo rows => so has b
    fdfd_info_t inf{};
    FDFD_TRY(S.solve(&inf));
    if (inf.flag != FDFD_OK) { fdfd_set_error(ctx, "fdfd_eigenfrequency_slab: inner solve %d failed (flag %d, relres %.2e)", j, inf.flag, inf.relres); return inf.flag; }
    inner_its += inf.iters; inner_ms += inf.solve_ms;
    // the solution's halo rows hold copies of the neighbours' rows: basis vectors keep zero halo rows
    CUDA_TRY(ctx, cudaMemsetAsync(wv.p, 0, Nl * sizeof(c128), st));
    CUDA_TRY(ctx, cudaMemcpyAsync(wv.p + S.H * S.Nx, S.w.x.p + S.H * S.Nx, own * sizeof(c128), cudaMemcpyDeviceToDevice, st));
    return FDFD_OK;
  };
  ops.dot_v_w = [&](int i, cd* h) -> int { return slab_dot(S, parts, V[i].p, wv.p, h); };
  ops.axpy_w = [&](int i, cd h) -> int { k_axpy_m<<<nb, 256, 0, st>>>(Nl, to_c128(h), V[i].p, wv.p); KLAUNCH(ctx); return FDFD_OK; };
  ops.norm_w = [&](double* nrm) -> int {
    cd n2;
    FDFD_TRY(slab_dot(S, parts, wv.p, wv.p, &n2));
    *nrm = std::sqrt(std::max(0.0, n2.real()));
    return FDFD_OK;
  };
  ops.set_v = [&](int j, double sc) -> int {
    FDFD_TRY(ensure(j));
    k_scale_real<<<nb, 256, 0, st>>>(Nl, sc, wv.p, V[j].p); KLAUNCH(ctx);
    return FDFD_OK;
  };
  ops.random_w = [&](int seed) -> int {   // pseudo-random on the owned rows (a function of the GLOBAL row), zero halo rows
    CUDA_TRY(ctx, cudaMemsetAsync(wv.p, 0, Nl * sizeof(c128), st));
    k_seed_rows<<<nb, 256, 0, st>>>(S.Nx, S.nyl, S.y0, wv.p + S.H * S.Nx, 20260101ull + 7919ull * (uint64_t)seed); KLAUNCH(ctx);
    return FDFD_OK;
  };
  ops.rotate_basis = [&](int m, int k, const std::vector<cd>& Q) -> int {
    std::vector<DevBuf<c128>> T(k);
    for (int i = 0; i < k; ++i) {
      if (T[i].alloc(Nl) != cudaSuccess) { cudaGetLastError(); fdfd_set_error(ctx, "out of device memory for the thick restart"); return FDFD_ERR_ALLOC; }
      for (int j = 0; j < m; ++j) { k_axpy_acc<<<nb, 256, 0, st>>>(Nl, to_c128(Q[(size_t)i * m + j]), V[j].p, T[i].p, j == 0); KLAUNCH(ctx); }
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    std::swap(V[k], V[m]);
    for (int i = 0; i < k; ++i) std::swap(V[i], T[i]);
    return FDFD_OK;
  };
  ArnoldiResult R;
  FDFD_TRY(krylov_schur(ctx, ops, nev, ncv, which, tol_eig, max_steps, o.verbose != 0, R));
  const int m = R.m;

  DevBuf<c128> ez, f3;
  CUDA_TRY(ctx, ez.alloc(Nl));
  for (int e = 0; e < nev; ++e) {
    const cd lam = sigma + 1.0 / R.nu[e];
    const cd om = std::sqrt(-lam / mu0 / eps0);   // eigen.jl:87
    omega_out[e].re = om.real(); omega_out[e].im = om.imag();
    if (!fields_rows) continue;
    for (int i = 0; i < m; ++i) { k_axpy_acc<<<nb, 256, 0, st>>>(Nl, to_c128(R.Y[(size_t)e * m + i]), V[i].p, ez.p, i == 0); KLAUNCH(ctx); }
    FDFD_TRY(S.fields_out(0, ez.p, 1, om, fields_rows + (size_t)e * 3 * own, f3));   // H from FORWARD differences (eigen.jl:90-91)
  }
  if (info) {
    std::memset(info, 0, sizeof(*info));
    info->iters = inner_its; info->flag = FDFD_OK; info->relres = o.tol; info->solve_ms = inner_ms;
    info->setup_ms = S.setup_ms; info->launches = ctx->launches - launches0; info->restarts = R.steps;   // restarts := Arnoldi steps (operator applications)
    info->mg_levels = S.ka >= 1 ? S.mgcs[0]->levels() : S.nlev;
    info->total_ms = wall_ms() - t0;
  }
  return FDFD_OK;
}

}  // namespace

extern "C" int fdfd_solve_modulated_slab(fdfd_ctx* ctx, fdfd_comm* comm, const fdfd_grid_t* g, double omega, double Omega,
                                         int nsidebands, int sharedpml, const fdfd_c128* eps_r_rows, const fdfd_c128* deps_r_rows,
                                         const fdfd_c128* src_rows, const fdfd_solve_opts_t* opts, fdfd_c128* fields_rows,
                                         fdfd_info_t* info) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  ARG_CHECK(ctx, comm != nullptr, "comm is NULL");
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, nsidebands >= 0 && nsidebands <= 16, "nsidebands out of range");
  ARG_CHECK(ctx, eps_r_rows && deps_r_rows && src_rows && fields_rows, "NULL argument");
  ARG_CHECK(ctx, omega > 0, "omega must be > 0");
  const int st = solve_modulated_slab(ctx, comm, g, omega, Omega, nsidebands, sharedpml, eps_r_rows, deps_r_rows, src_rows, opts,
                                      fields_rows, info);
  // a Krylov flag (no convergence) is reported on every rank alike and leaves the communicator usable
  if (st != FDFD_OK && st != FDFD_ERR_NOCONV && st != FDFD_ERR_BREAKDOWN) release_group(comm, st);
  return st;
}

extern "C" int fdfd_eigenfrequency_slab(fdfd_ctx* ctx, fdfd_comm* comm, const fdfd_grid_t* g, int pol, double omega0, int nev,
                                        int which, int ncv, const fdfd_c128* eps_r_rows, const fdfd_solve_opts_t* opts,
                                        fdfd_c128* omega_out, fdfd_c128* fields_rows, fdfd_info_t* info) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  ARG_CHECK(ctx, comm != nullptr, "comm is NULL");
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, pol == FDFD_TM, "the slab eigenfrequency solve is TM only (TE slabs are not built)");
  ARG_CHECK(ctx, nev >= 1 && eps_r_rows && omega_out, "bad arguments");
  ARG_CHECK(ctx, which >= FDFD_WHICH_LM && which <= FDFD_WHICH_SI, "bad `which`");
  ARG_CHECK(ctx, omega0 > 0, "omega0 must be > 0");
  const int st = eigenfrequency_slab(ctx, comm, g, omega0, nev, which, ncv, eps_r_rows, opts, omega_out, fields_rows, info);
  if (st != FDFD_OK && st != FDFD_ERR_NOCONV && st != FDFD_ERR_BREAKDOWN) release_group(comm, st);
  return st;
}
