// api.cu -- C-ABI entry points built on the resident problem handle: driven TM/TE solve.
#include "krylov.cuh"
#include <atomic>
#include <chrono>
#include <cmath>
#include <thread>

namespace {

__global__ void k_scale_src(int64_t N, c128 k, const c128* __restrict__ src, c128* __restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    b[i] = k * src[i];
}

__global__ void k_fill_random(int64_t N, c128* __restrict__ x, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    x[i] = c128((double)(z & 0xFFFFFFFF) / 4294967296.0 - 0.5, (double)(z >> 32) / 4294967296.0 - 0.5);
  }
}

template <typename T> __global__ void k_cast_in(int64_t N, const c128* __restrict__ in, cplx<T>* __restrict__ out, double sc) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) out[i] = cplx<T>(c128(sc * in[i].x, sc * in[i].y));
}
template <typename T> __global__ void k_cast_out(int64_t N, const cplx<T>* __restrict__ in, c128* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) out[i] = c128(in[i]);
}

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int vec_blocks(fdfd_ctx* ctx, int64_t N) { return (int)std::min<int64_t>((N + 255) / 256, (int64_t)ctx->num_sms * 8); }

}  // namespace

MGParams mg_params_from(const fdfd_solve_opts_t& o) {
  MGParams m;
  m.cycle = o.mg_cycle; m.wdepth = o.mg_wdepth; m.nu1 = std::max(1, o.mg_nu1); m.nu2 = std::max(0, o.mg_nu2);
  m.coarse_sweeps = std::max(1, o.mg_coarse_sweeps);
  m.beta = o.mg_beta; m.wjac = o.mg_wjac; m.wline = o.mg_wline;
  m.shift_growth = o.mg_shift_growth; if (o.mg_max_levels > 0) m.max_levels = o.mg_max_levels;
  if (const char* e = getenv("FDFD_MG_PAD")) m.pad = atoi(e);    // diagnostics only
  if (const char* e = getenv("FDFD_MG_MINN")) m.min_n = atoi(e);
  if (const char* e = getenv("FDFD_MG_KHSTOP")) m.kh_stop = atof(e);
  return m;
}

extern "C" int fdfd_problem_create(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, int ordering, double omega,
                                   const fdfd_c128* eps_r, const fdfd_solve_opts_t* opts, fdfd_problem** out) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  ARG_CHECK(ctx, out != nullptr, "out is NULL");
  *out = nullptr;
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, pol == FDFD_TM || pol == FDFD_TE, "pol must be FDFD_TM or FDFD_TE");
  ARG_CHECK(ctx, ordering == FDFD_ORDER_FB || ordering == FDFD_ORDER_BF, "bad ordering");
  ARG_CHECK(ctx, eps_r != nullptr, "eps_r is NULL");
  ARG_CHECK(ctx, omega > 0, "omega must be > 0");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  fdfd_problem* P = new fdfd_problem();
  P->ctx = ctx;
  if (opts) P->opts = *opts; else fdfd_default_opts(&P->opts);
  if (P->opts.solver == FDFD_SOLVER_AUTO) {
    // measured crossover: below ~2048^2 both solvers are launch-latency bound and BiCGSTAB is as fast
    const bool big = g->Nx * g->Ny >= ((int64_t)1 << 22) && std::min(g->Nx, g->Ny) >= 1024;
    P->opts.solver = (big && P->opts.precond == FDFD_PRECOND_MG && P->opts.mg_precision == FDFD_MG_F32) ? FDFD_SOLVER_MLKRYLOV : FDFD_SOLVER_BICGSTAB;
  }
  if (!(P->opts.solver == FDFD_SOLVER_BICGSTAB || P->opts.solver == FDFD_SOLVER_COCG || P->opts.solver == FDFD_SOLVER_MLKRYLOV)) { delete P; fdfd_set_error(ctx, "fdfd_problem_create: solver must be FDFD_SOLVER_BICGSTAB, FDFD_SOLVER_COCG or FDFD_SOLVER_MLKRYLOV"); return FDFD_ERR_ARG; }
  if (P->opts.solver == FDFD_SOLVER_MLKRYLOV && !(P->opts.precond == FDFD_PRECOND_MG && P->opts.mg_precision == FDFD_MG_F32)) { delete P; fdfd_set_error(ctx, "fdfd_problem_create: FDFD_SOLVER_MLKRYLOV needs FDFD_PRECOND_MG and FDFD_MG_F32"); return FDFD_ERR_ARG; }
  if (P->opts.solver == FDFD_SOLVER_COCG && P->opts.precond == FDFD_PRECOND_MG) { delete P; fdfd_set_error(ctx, "fdfd_problem_create: COCG needs a symmetric preconditioner (FDFD_PRECOND_JACOBI or FDFD_PRECOND_NONE); the multigrid cycle is not symmetric"); return FDFD_ERR_ARG; }
  const double t0 = now_ms();
  int st = P->op.build(ctx, *g, pol, ordering, omega, eps_r);
  if (st != FDFD_OK) { delete P; return st; }
  const int64_t N = g->Nx * g->Ny;
  auto fail = [&](int code) { delete P; return code; };
  st = P->w.alloc(ctx, N, apply_num_blocks(g->Nx, g->Ny), P->opts.maxit, P->opts.precond == FDFD_PRECOND_JACOBI);
  if (st != FDFD_OK) return fail(st);
  if (P->opts.precond == FDFD_PRECOND_MG) {
    MGParams mp = mg_params_from(P->opts);
    if (P->opts.mg_precision == FDFD_MG_F64) { P->mgd = new Multigrid<double>(); st = P->mgd->setup(ctx, P->op, mp); }
    else { P->mgf = new Multigrid<float>(); st = P->mgf->setup(ctx, P->op, mp); }
    if (st != FDFD_OK) return fail(st);
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { fdfd_set_error(ctx, "setup failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(FDFD_ERR_CUDA); }
  P->setup_ms = now_ms() - t0;
  *out = P;
  return FDFD_OK;
}

extern "C" void fdfd_problem_destroy(fdfd_problem* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  delete p;
}

extern "C" int fdfd_problem_set_rhs(fdfd_problem* P, const fdfd_c128* b) {
  if (!P) return FDFD_ERR_ARG;
  ARG_CHECK(P->ctx, b != nullptr, "b is NULL");
  const int64_t N = P->op.g.Nx * P->op.g.Ny;
  FDFD_TRY(fdfd_copy_in(P->ctx, P->w.b.p, b, N * sizeof(c128)));
  CUDA_TRY(P->ctx, cudaStreamSynchronize(P->ctx->stream));
  P->have_rhs = true;
  return FDFD_OK;
}

extern "C" int fdfd_problem_set_source(fdfd_problem* P, const fdfd_c128* src) {
  if (!P) return FDFD_ERR_ARG;
  fdfd_ctx* ctx = P->ctx;
  ARG_CHECK(ctx, src != nullptr, "src is NULL");
  const int64_t N = P->op.g.Nx * P->op.g.Ny;
  // stage src in t (scratch), b = 1im*ω*src (driven.jl:36)
  FDFD_TRY(fdfd_copy_in(ctx, P->w.t.p, src, N * sizeof(c128)));
  k_scale_src<<<P->w.nvec_blocks, 256, 0, ctx->stream>>>(N, c128(0.0, P->op.omega), P->w.t.p, P->w.b.p); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  P->have_rhs = true;
  return FDFD_OK;
}

extern "C" int fdfd_problem_solve(fdfd_problem* P, fdfd_info_t* info) {
  if (!P) return FDFD_ERR_ARG;
  ARG_CHECK(P->ctx, P->have_rhs, "no right-hand side set");
  fdfd_info_t local{};
  if (!info) info = &local;
  std::memset(info, 0, sizeof(*info));
  CUDA_TRY(P->ctx, cudaSetDevice(P->ctx->device));
  const double t0 = now_ms();
  if (P->opts.solver == FDFD_SOLVER_COCG) { FDFD_TRY(krylov_cocg(P, info)); }
  else if (P->opts.solver == FDFD_SOLVER_MLKRYLOV) { FDFD_TRY(krylov_multilevel(P, info)); }
  else { KrylovOps ops = P->make_ops(); FDFD_TRY(krylov_bicgstab(P->ctx, P->w, ops, P->opts, info)); }
  info->setup_ms = P->setup_ms;
  info->mg_levels = P->mgf ? P->mgf->levels() : (P->mgd ? P->mgd->levels() : 0);
  info->total_ms = now_ms() - t0;
  if (info->flag != FDFD_OK) fdfd_set_error(P->ctx, "Krylov solver stopped with flag %d after %d iterations, relres %.3e", info->flag, info->iters, info->relres);
  return FDFD_OK;  // convergence state is reported through info->flag; results are valid approximations
}

extern "C" int fdfd_problem_get_solution(fdfd_problem* P, fdfd_c128* x) {
  if (!P) return FDFD_ERR_ARG;
  ARG_CHECK(P->ctx, x != nullptr, "x is NULL");
  const int64_t N = P->op.g.Nx * P->op.g.Ny;
  FDFD_TRY(fdfd_copy_out(P->ctx, x, P->w.x.p, N * sizeof(c128)));
  CUDA_TRY(P->ctx, cudaStreamSynchronize(P->ctx->stream));
  return FDFD_OK;
}

extern "C" int fdfd_problem_get_fields(fdfd_problem* P, int forward_h, fdfd_c128* fields) {
  if (!P) return FDFD_ERR_ARG;
  fdfd_ctx* ctx = P->ctx;
  ARG_CHECK(ctx, fields != nullptr, "fields is NULL");
  const int64_t N = P->op.g.Nx * P->op.g.Ny;
  DevBuf<c128> f3;
  CUDA_TRY(ctx, f3.alloc(3 * N));
  FDFD_TRY(launch_recover(ctx, P->op, P->w.x.p, forward_h, std::complex<double>(P->op.omega, 0.0), 0, f3.p));
  FDFD_TRY(fdfd_copy_out(ctx, fields, f3.p, 3 * N * sizeof(c128)));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FDFD_OK;
}

extern "C" int fdfd_problem_bench_apply(fdfd_problem* P, int nrep, double* ms_per_apply) {
  if (!P) return FDFD_ERR_ARG;
  fdfd_ctx* ctx = P->ctx;
  ARG_CHECK(ctx, nrep > 0 && ms_per_apply, "bad arguments");
  const int64_t N = P->op.g.Nx * P->op.g.Ny;
  const bool te = P->op.pol == FDFD_TE;
  k_fill_random<<<P->w.nvec_blocks, 256, 0, ctx->stream>>>(N, P->w.p.p, 1234); KLAUNCH(ctx);
  DotSpec ds;
  for (int w = 0; w < 3; ++w) FDFD_TRY(launch_apply(ctx, P->op.view(), te, P->w.p.p, false, P->w.v.p, ds));
  cudaEvent_t e0, e1;
  CUDA_TRY(ctx, cudaEventCreate(&e0)); CUDA_TRY(ctx, cudaEventCreate(&e1));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
  for (int i = 0; i < nrep; ++i) {
    // ping-pong so consecutive launches do not hit identical cache state
    FDFD_TRY(launch_apply(ctx, P->op.view(), te, (i & 1) ? P->w.v.p : P->w.p.p, false, (i & 1) ? P->w.p.p : P->w.v.p, ds));
  }
  CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
  CUDA_TRY(ctx, cudaEventSynchronize(e1));
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_per_apply = (double)ms / nrep;
  return FDFD_OK;
}

// ms per launch of the batched stencil (nrhs right-hand sides sharing the problem's operator), timed like fdfd_problem_bench_apply
extern "C" int fdfd_problem_bench_apply_batched(fdfd_problem* P, int nrhs, int nrep, double* ms_per_launch) {
  if (!P) return FDFD_ERR_ARG;
  fdfd_ctx* ctx = P->ctx;
  ARG_CHECK(ctx, nrep > 0 && ms_per_launch && (nrhs == 1 || nrhs == 2 || nrhs == 4 || nrhs == 8), "bad arguments (nrhs in {1, 2, 4, 8})");
  const int64_t N = P->op.g.Nx * P->op.g.Ny;
  const bool te = P->op.pol == FDFD_TE;
  DevBuf<c128> xa, xb;
  CUDA_TRY(ctx, xa.alloc((size_t)nrhs * N)); CUDA_TRY(ctx, xb.alloc((size_t)nrhs * N));
  k_fill_random<<<P->w.nvec_blocks, 256, 0, ctx->stream>>>((int64_t)nrhs * N, xa.p, 99); KLAUNCH(ctx);
  for (int w = 0; w < 3; ++w) FDFD_TRY(launch_apply_batched(ctx, P->op.view(), te, xa.p, xb.p, nrhs, N));
  cudaEvent_t e0, e1;
  CUDA_TRY(ctx, cudaEventCreate(&e0)); CUDA_TRY(ctx, cudaEventCreate(&e1));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
  for (int i = 0; i < nrep; ++i) FDFD_TRY(launch_apply_batched(ctx, P->op.view(), te, (i & 1) ? xb.p : xa.p, (i & 1) ? xa.p : xb.p, nrhs, N));
  CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
  CUDA_TRY(ctx, cudaEventSynchronize(e1));
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_per_launch = (double)ms / nrep;
  return FDFD_OK;
}

extern "C" int fdfd_problem_bench_mg(fdfd_problem* P, int kind, int nrep, double* ms_per_launch) {
  if (!P) return FDFD_ERR_ARG;
  fdfd_ctx* ctx = P->ctx;
  ARG_CHECK(ctx, nrep > 0 && ms_per_launch, "bad arguments");
  ARG_CHECK(ctx, P->mgf != nullptr && P->mgf->levels() >= 2, "needs the fp32 multigrid with at least two levels");
  Multigrid<float>* mg = P->mgf;
  const int64_t N = P->op.g.Nx * P->op.g.Ny;
  const int* saved = mg->done; mg->done = nullptr;
  // a realistic state: random right-hand side, one full cycle so that every level holds an iterate
  k_fill_random<<<P->w.nvec_blocks, 256, 0, ctx->stream>>>(N, P->w.t.p, 4321); KLAUNCH(ctx);
  k_cast_in<float><<<P->w.nvec_blocks, 256, 0, ctx->stream>>>(N, P->w.t.p, mg->rhs(), 1.0); KLAUNCH(ctx);
  const c64* res = nullptr;
  int st = mg->apply(&res);
  auto one = [&]() -> int {
    switch (kind) {
      case 0: return mg->smooth(0, false, true);
      case 1: return mg->restrict_residual(0);
      case 2: return mg->smooth(0, true, false);
      default: return mg->apply(&res);
    }
  };
  for (int w = 0; w < 3 && st == FDFD_OK; ++w) st = one();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (st == FDFD_OK && (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess)) st = FDFD_ERR_CUDA;
  float ms = 0;
  if (st == FDFD_OK) {
    cudaStreamSynchronize(ctx->stream);
    cudaEventRecord(e0, ctx->stream);
    for (int i = 0; i < nrep && st == FDFD_OK; ++i) st = one();
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  mg->done = saved;
  FDFD_TRY(st);
  *ms_per_launch = (double)ms / nrep;
  return FDFD_OK;
}

extern "C" int fdfd_problem_get_history(fdfd_problem* P, double* out, int n, int* written) {
  if (!P) return FDFD_ERR_ARG;
  fdfd_ctx* ctx = P->ctx;
  ARG_CHECK(ctx, out && n > 0 && written, "bad arguments");
  const KScal h = *P->w.h_scal;
  const int cnt = std::max(0, std::min(std::min(n, h.iter + 1), (int)P->w.hist.n));
  std::vector<double> tmp(cnt);
  CUDA_TRY(ctx, cudaMemcpyAsync(tmp.data(), P->w.hist.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < cnt; ++k) out[k] = k == 0 ? 1.0 : std::sqrt(tmp[k] / h.bnorm2);
  *written = cnt;
  return FDFD_OK;
}

extern "C" int fdfd_problem_precond(fdfd_problem* P, const fdfd_c128* in, fdfd_c128* out) {
  if (!P) return FDFD_ERR_ARG;
  fdfd_ctx* ctx = P->ctx;
  ARG_CHECK(ctx, in && out, "NULL argument");
  ARG_CHECK(ctx, P->mgf || P->mgd, "problem has no multigrid preconditioner");
  const int64_t N = P->op.g.Nx * P->op.g.Ny;
  FDFD_TRY(fdfd_copy_in(ctx, P->w.t.p, in, N * sizeof(c128)));
  const int blocks = P->w.nvec_blocks;
  if (P->mgf) {
    const int* saved = P->mgf->done; P->mgf->done = nullptr;
    k_cast_in<float><<<blocks, 256, 0, ctx->stream>>>(N, P->w.t.p, P->mgf->rhs(), P->mgf->rhs_scale); KLAUNCH(ctx);
    const c64* res = nullptr;
    int st = P->mgf->apply(&res);
    P->mgf->done = saved;
    FDFD_TRY(st);
    k_cast_out<float><<<blocks, 256, 0, ctx->stream>>>(N, res, P->w.t.p); KLAUNCH(ctx);
  } else {
    const int* saved = P->mgd->done; P->mgd->done = nullptr;
    k_cast_in<double><<<blocks, 256, 0, ctx->stream>>>(N, P->w.t.p, P->mgd->rhs(), P->mgd->rhs_scale); KLAUNCH(ctx);
    const c128* res = nullptr;
    int st = P->mgd->apply(&res);
    P->mgd->done = saved;
    FDFD_TRY(st);
    k_cast_out<double><<<blocks, 256, 0, ctx->stream>>>(N, res, P->w.t.p); KLAUNCH(ctx);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  FDFD_TRY(fdfd_copy_out(ctx, out, P->w.t.p, N * sizeof(c128)));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FDFD_OK;
}

// ---- solve(d::Device, pol) ------------------------------------------------------------------------------
// The reference sweeps its frequencies serially (`for i in eachindex(d.ω)`, driven.jl:11) and rebuilds everything per
// ω.  The units are independent, and one solve leaves the GPU latency-bound on its coarse multigrid levels, so up to
// `opts->concurrency` frequencies are solved at the same time, each on its own stream / CUDA-graph (worker threads
// pull the next frequency from a shared counter).  Measured at 4096^2: 4 concurrent solves -> 1.7x the throughput.
static int solve_one_omega(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, double omega, const fdfd_c128* eps_r,
                           const fdfd_c128* src, const fdfd_solve_opts_t* opts, fdfd_c128* fields, fdfd_info_t* inf) {
  const double t0 = now_ms();
  fdfd_problem* P = nullptr;
  FDFD_TRY(fdfd_problem_create(ctx, g, pol, FDFD_ORDER_FB, omega, eps_r, opts, &P));
  int st = fdfd_problem_set_source(P, src);
  if (st == FDFD_OK) st = fdfd_problem_solve(P, inf);
  // driven.jl:40-41 (TM, backward) / 50-51 (TE, backward)
  if (st == FDFD_OK) st = fdfd_problem_get_fields(P, 0, fields);
  fdfd_problem_destroy(P);
  inf->total_ms = now_ms() - t0;
  return st;
}

extern "C" int fdfd_solve_driven(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, int n_omega, const double* omega,
                                 const fdfd_c128* eps_r, const fdfd_c128* src, int src_per_omega,
                                 const fdfd_solve_opts_t* opts, fdfd_c128* fields, fdfd_info_t* info) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, n_omega >= 1 && omega, "need at least one frequency");
  ARG_CHECK(ctx, eps_r && src && fields, "NULL argument");
  const int64_t N = g->Nx * g->Ny;
  fdfd_solve_opts_t o;
  if (opts) o = *opts; else fdfd_default_opts(&o);
  const int nworkers = std::max(1, std::min(n_omega, o.concurrency > 0 ? o.concurrency : 1));
  if (((uint32_t)o.ml_spec >> 24) == 0) {
    // multilevel Krylov: the level-0 basis (2 vectors per outer iteration) of every worker gets an equal share of the free device
    // memory, so that no worker starves the others into short restart cycles
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t fr = fdfd_dev_mem_available();
    if (fr > 0) {
      const double per = 0.7 * (double)fr / nworkers - 24.0 * 16.0 * (double)N;   // minus Krylov vectors, hierarchy, inner levels, fields
      const int r = (int)std::max(16.0, std::min(96.0, per / (2.0 * 16.0 * (double)N)));
      o.ml_spec = (int32_t)(((uint32_t)o.ml_spec & 0xffffffu) | ((uint32_t)r << 24));
    }
  }
  std::vector<fdfd_info_t> infos(n_omega);
  std::vector<int> status(n_omega, FDFD_OK);
  // host inputs are staged into HBM ONCE per call (the reference re-reads d.eps_r / d.src for every frequency, driven.jl:21,36):
  // every frequency's problem then copies device to device, and the call moves 2 N (or (1 + n_omega) N) complex numbers over
  // PCIe instead of 2 N n_omega.  The staging runs on the caller's stream, which the worker streams are ordered against below.
  DevBuf<c128> eps_stage, src_stage;
  if (n_omega > 1) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!fdfd_is_device_ptr(eps_r)) {
      CUDA_TRY(ctx, eps_stage.alloc(N));
      FDFD_TRY(fdfd_copy_in(ctx, eps_stage.p, eps_r, N * sizeof(c128)));
      eps_r = reinterpret_cast<const fdfd_c128*>(eps_stage.p);
    }
    if (!fdfd_is_device_ptr(src)) {
      const size_t cnt = (size_t)N * (src_per_omega ? n_omega : 1);
      CUDA_TRY(ctx, src_stage.alloc(cnt));
      FDFD_TRY(fdfd_copy_in(ctx, src_stage.p, src, cnt * sizeof(c128)));
      src = reinterpret_cast<const fdfd_c128*>(src_stage.p);
    }
    // workers run on private streams: the caller's stream (which may carry the producers of device-resident eps_r / src, and
    // carries the staging copies above) must be drained before they start; every worker synchronises its own stream before it
    // returns, so the outputs are complete when the call returns
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (nworkers == 1) {
    for (int i = 0; i < n_omega; ++i) {
      status[i] = solve_one_omega(ctx, g, pol, omega[i], eps_r, src + (src_per_omega ? (size_t)i * N : 0), &o,
                                  fields + (size_t)i * 3 * N, &infos[i]);
      if (status[i] != FDFD_OK) return status[i];
    }
  } else {
    std::atomic<int> next(0);
    std::vector<std::string> errs(nworkers);
    std::vector<int64_t> launches(nworkers, 0);
    std::vector<std::thread> th;
    for (int wkr = 0; wkr < nworkers; ++wkr) {
      th.emplace_back([&, wkr]() {
        fdfd_ctx* sub = nullptr;
        if (fdfd_ctx_create(ctx->device, nullptr, &sub) != FDFD_OK) { errs[wkr] = "fdfd_solve_driven: could not create a worker context (stream)"; return; }
        for (int i = next.fetch_add(1); i < n_omega; i = next.fetch_add(1)) {
          status[i] = solve_one_omega(sub, g, pol, omega[i], eps_r, src + (src_per_omega ? (size_t)i * N : 0), &o,
                                      fields + (size_t)i * 3 * N, &infos[i]);
          if (status[i] != FDFD_OK) { errs[wkr] = sub->err; break; }
        }
        launches[wkr] = sub->launches;
        fdfd_ctx_destroy(sub);
      });
    }
    for (auto& t : th) t.join();
    for (int wkr = 0; wkr < nworkers; ++wkr) ctx->launches += launches[wkr];
    for (int i = 0; i < n_omega; ++i)
      if (status[i] != FDFD_OK) {
        for (auto& e : errs) if (!e.empty()) { fdfd_set_error(ctx, "%s", e.c_str()); break; }
        return status[i];
      }
  }
  int worst = FDFD_OK;
  for (int i = 0; i < n_omega; ++i) {
    if (info) info[i] = infos[i];
    if (infos[i].flag != FDFD_OK) worst = infos[i].flag;
  }
  if (worst != FDFD_OK) { fdfd_set_error(ctx, "fdfd_solve_driven: at least one frequency did not converge (flag %d)", worst); return worst; }
  return FDFD_OK;
}
