// modulated.cu -- K6: the multi-frequency (MF-FDFD) solve of a time-modulated device.
// Replaces solve(d::ModulatedDevice) (src/solver/modulation.jl:35-119): nf = 2 ns + 1 sidebands w_n = w + n W,
//   (A1_n + w_n^2 Teps) e_n + 1/2 w_n^2 conj(TDeps) e_{n+1} + 1/2 w_n^2 TDeps e_{n-1} = i w src delta_{n0}
// with A1 in the b.f ordering (modulation.jl:82) and the PML evaluated at w (sharedpml) or at w_n (:87-91).
// The block system is never assembled: one stencil launch per sideband carries the pointwise coupling, the
// sidebands' vectors are slices of one (nf*N) Krylov vector, and the preconditioner is block diagonal (one
// shifted-Laplacian multigrid per sideband; the coupling, |Deps| << eps, is left to the Krylov iteration).
// H is recovered with FORWARD stretched differences at the sideband's own frequency (modulation.jl:112-113).
#include "krylov.cuh"
#include <chrono>
#include <memory>

namespace {

__global__ void k_mod_rhs(int64_t N, int nf, int centre, c128 k, const c128* __restrict__ src, c128* __restrict__ b) {
  const int64_t tot = N * nf;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / N, n = i - j * N;
    b[i] = (j == centre) ? k * src[n] : c128(0.0, 0.0);
  }
}

template <typename T> struct ModSystem {
  fdfd_ctx* ctx = nullptr;
  int nf = 0;
  int64_t N = 0;
  std::vector<std::unique_ptr<FineOp>> ops;
  std::vector<std::unique_ptr<Multigrid<T>>> mgs;
  std::vector<double> omegan;
  DevBuf<c128> deps;            // eps0-free Deps_r as given
  DevBuf<cplx<T>> F, U, Tm, S;  // contiguous level-0 buffers of all sidebands (rhs, u, tmp, spare)
  double eps0 = 0;
  KrylovWork w;

  int apply(const void* x, bool x_f32, c128* y, const DotSpec& ds) {
    const int nab = apply_num_blocks(ops[0]->g.Nx, ops[0]->g.Ny);
    const size_t esz = x_f32 ? sizeof(c64) : sizeof(c128);
    for (int j = 0; j < nf; ++j) {
      DotSpec d = ds;
      if (ds.ndot > 0) { d.partials = ds.partials + (size_t)j * nab * ds.ndot; d.d0 = ds.d0 + (size_t)j * N; }
      Coupling c;
      c.deps = deps.p;
      c.hw = 0.5 * omegan[j] * omegan[j] * eps0;  // 0.5*ωn[j]^2 * (ϵ₀ L₀)  (modulation.jl:95-98)
      c.xm1 = j > 0 ? (const char*)x + (size_t)(j - 1) * N * esz : nullptr;
      c.xp1 = j + 1 < nf ? (const char*)x + (size_t)(j + 1) * N * esz : nullptr;
      FDFD_TRY(launch_apply(ctx, ops[j]->view(), false, (const char*)x + (size_t)j * N * esz, x_f32, y + (size_t)j * N, d, &c));
    }
    return FDFD_OK;
  }

  int precond(bool hold, const void** out) {
    const cplx<T>* r0 = nullptr;
    for (int j = 0; j < nf; ++j) {
      const cplx<T>* res = nullptr;
      FDFD_TRY(mgs[j]->apply(&res));
      if (hold) { std::swap(mgs[j]->lv[0].u.p, mgs[j]->spare.p); res = mgs[j]->spare.p; }
      if (j == 0) r0 = res;
    }
    *out = r0;  // sideband buffers rotate in lock step, so the results stay contiguous
    return FDFD_OK;
  }
};

template <typename T>
int solve_modulated_t(fdfd_ctx* ctx, const fdfd_grid_t* g, double omega, double Omega, int ns, int sharedpml,
                      const fdfd_c128* eps_r, const fdfd_c128* deps_r, const fdfd_c128* src, const fdfd_solve_opts_t& o,
                      fdfd_c128* fields, fdfd_info_t* info) {
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  const int nf = 2 * ns + 1;
  const int64_t N = g->Nx * g->Ny;
  ModSystem<T> M;
  M.ctx = ctx; M.nf = nf; M.N = N; M.eps0 = kEps0 * g->L0;
  MGParams mp = mg_params_from(o);
  CUDA_TRY(ctx, M.deps.alloc(N));
  FDFD_TRY(fdfd_copy_in(ctx, M.deps.p, deps_r, N * sizeof(c128)));
  CUDA_TRY(ctx, M.F.alloc((size_t)nf * N)); CUDA_TRY(ctx, M.U.alloc((size_t)nf * N));
  CUDA_TRY(ctx, M.Tm.alloc((size_t)nf * N)); CUDA_TRY(ctx, M.S.alloc((size_t)nf * N));
  FDFD_TRY(M.w.alloc(ctx, (int64_t)nf * N, nf * apply_num_blocks(g->Nx, g->Ny), o.maxit, false));
  for (int j = 0; j < nf; ++j) {
    const double wn = omega + Omega * (double)(j - ns);  // ωn = ω .+ Ω*n  (modulation.jl:41,49)
    ARG_CHECK(ctx, wn > 0, "a sideband frequency w + n*Omega is not positive");
    M.omegan.push_back(wn);
    M.ops.emplace_back(new FineOp());
    FDFD_TRY(M.ops[j]->build(ctx, *g, FDFD_TM, FDFD_ORDER_BF, wn, eps_r, sharedpml ? omega : wn));
    M.mgs.emplace_back(new Multigrid<T>());
    FDFD_TRY(M.mgs[j]->setup(ctx, *M.ops[j], mp));
    MGLevel<T>& L0 = M.mgs[j]->lv[0];
    L0.f.alias(M.F.p + (size_t)j * N, N); L0.u.alias(M.U.p + (size_t)j * N, N); L0.tmp.alias(M.Tm.p + (size_t)j * N, N);
    M.mgs[j]->spare.alias(M.S.p + (size_t)j * N, N);
  }
  // b: zeros(N*nf); centre block = 1im*ω*src  (modulation.jl:67-69)
  FDFD_TRY(fdfd_copy_in(ctx, M.w.t.p, src, N * sizeof(c128)));
  k_mod_rhs<<<M.w.nvec_blocks, 256, 0, ctx->stream>>>(N, nf, ns, c128(0.0, omega), M.w.t.p, M.w.b.p); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const double setup_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();

  KrylovOps k;
  k.prec_f32 = sizeof(T) == sizeof(float);
  k.prec_rhs = M.F.p; k.fscale = M.mgs[0]->rhs_scale;
  k.nab = nf * apply_num_blocks(g->Nx, g->Ny);
  k.apply = [&M](const void* x, bool x_f32, c128* y, const DotSpec& ds) { return M.apply(x, x_f32, y, ds); };
  k.precond = [&M](bool hold, const void** out) { return M.precond(hold, out); };
  k.get_state = [&M](std::vector<void*>& v) {
    v.clear();
    for (auto& mg : M.mgs) { for (auto& L : mg->lv) { v.push_back(L.u.p); v.push_back(L.tmp.p); } v.push_back(mg->spare.p); }
  };
  k.set_state = [&M](const std::vector<void*>& v) {
    size_t i = 0;
    for (auto& mg : M.mgs) { for (auto& L : mg->lv) { L.u.p = (cplx<T>*)v[i++]; L.tmp.p = (cplx<T>*)v[i++]; } mg->spare.p = (cplx<T>*)v[i++]; }
  };
  fdfd_info_t inf{};
  FDFD_TRY(krylov_bicgstab(ctx, M.w, k, o, &inf));
  inf.setup_ms = setup_ms;
  inf.mg_levels = M.mgs[0]->levels();
  // fields per sideband: hx = -1/1im/ωn/μ₀*Syf*δyf*ez, hy = 1/1im/ωn/μ₀*Sxf*δxf*ez
  DevBuf<c128> f3;
  CUDA_TRY(ctx, f3.alloc(3 * N));
  for (int j = 0; j < nf; ++j) {
    FDFD_TRY(launch_recover(ctx, *M.ops[j], M.w.x.p + (size_t)j * N, 1, std::complex<double>(M.omegan[j], 0.0), 0, f3.p));
    FDFD_TRY(fdfd_copy_out(ctx, fields + (size_t)j * 3 * N, f3.p, 3 * N * sizeof(c128)));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  inf.total_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  if (info) *info = inf;
  if (inf.flag != FDFD_OK) {
    fdfd_set_error(ctx, "fdfd_solve_modulated: Krylov solver stopped with flag %d after %d iterations, relres %.3e", inf.flag, inf.iters, inf.relres);
    return inf.flag;
  }
  return FDFD_OK;
}

}  // namespace

extern "C" int fdfd_solve_modulated(fdfd_ctx* ctx, const fdfd_grid_t* g, double omega, double Omega, int nsidebands,
                                    int sharedpml, const fdfd_c128* eps_r, const fdfd_c128* deps_r, const fdfd_c128* src,
                                    const fdfd_solve_opts_t* opts, fdfd_c128* fields, fdfd_info_t* info) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, nsidebands >= 0 && nsidebands <= 16, "nsidebands out of range");
  ARG_CHECK(ctx, eps_r && deps_r && src && fields, "NULL argument");
  ARG_CHECK(ctx, omega > 0, "omega must be > 0");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  fdfd_solve_opts_t o;
  if (opts) o = *opts; else fdfd_default_opts(&o);
  ARG_CHECK(ctx, (o.solver == FDFD_SOLVER_BICGSTAB || o.solver == FDFD_SOLVER_AUTO) && o.precond == FDFD_PRECOND_MG, "modulated solve needs BiCGSTAB + multigrid");
  if (o.mg_precision == FDFD_MG_F64)
    return solve_modulated_t<double>(ctx, g, omega, Omega, nsidebands, sharedpml, eps_r, deps_r, src, o, fields, info);
  return solve_modulated_t<float>(ctx, g, omega, Omega, nsidebands, sharedpml, eps_r, deps_r, src, o, fields, info);
}
