// linsolve.cu -- dolinearsolve-level entry (SURVEY §8b "optional dolinearsolve-level export taking CSC", §8f row 4):
// x = A^-1 b for an ASSEMBLED SparseMatrixCSC handed over by the caller, for the reference call sites that build their own
// matrix and only use the solver seam -- dolinearsolve(A, b, matrixsym) in src/solver/solver.jl:4-41, called from
// src/solver/nonlinear.jl:69,97,120 (chi-3 outer loops) and usable by src/solver/eigen.jl:32-66.
//
// Not the headline path: the driven / modulated / eigenfrequency entry points are matrix-free and multigrid-preconditioned.
// A general sparse matrix carries no grid, so this path is the operator-agnostic BiCGSTAB of krylov.cu with
//   * apply  = SpMV over a SELL-32 layout (sliced ELLPACK, slice = one warp of rows, entries of a slice stored k-major so
//     that the 32 lanes of a warp read 32 consecutive values / column indices: HBM-coalesced without a segmented reduction;
//     int32 column indices; algorithmic bytes per stored entry 16 + 4, plus 16 B/row for y, 16 B/row gathered x from L2),
//     the Krylov dot products fused into the SpMV exactly like in k_apply (stencil.cu),
//   * precond = diagonal (Jacobi) scaling in fp64.
// The CSC -> SELL transposition is host work inside the call (the caller's arrays are host arrays; O(nnz), one pass to count,
// one to scatter) and its arithmetic core is shared with the host-only test hook fdfd_debug_sell_spmv.
//
// Measured on a B200 (round 2, tools/gpu_linsolve.py): k_sell_spmv 6.4-7.2 TB/s (212 B/row, 0.97-1.09 of the measured copy
// bandwidth; the gathered x comes from L2), grid-hinted seam = solve(d) in iterations and time.  GPU tests:
// tests/test_gpu_dolinearsolve.py; the host transposition / SELL indexing is also checked on the CPU (tests/test_cabi_cpu.py).
#include "krylov.cuh"
#include "reduce.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>

namespace {

constexpr int kSlice = 32;       // rows per slice = lanes per warp
constexpr int kSpThreads = 256;  // 8 slices per CTA

// host-side SELL-32 image of a CSC matrix
struct SellHost {
  int64_t n = 0, nslices = 0;
  std::vector<int64_t> sptr;     // nslices + 1: first entry of slice s (entries of a slice: width_s * 32, k-major)
  std::vector<int32_t> col;      // padded entries point at the row itself with value 0
  std::vector<c128> val;
  std::vector<c128> dinv;        // 1 / A[i,i]  (1 where the diagonal is absent or zero)
  std::vector<c128> rowsum;      // sum_j A[i,j] = (A 1)[i]: the mass term of a difference stencil (grid-hinted path)
  int64_t nnz = 0;
};

// CSC (colptr n+1, rowval, nzval; indices with the given base; duplicates are summed like Julia's sparse(I,J,V)) -> SELL-32
int csc_to_sell(int64_t n, const int64_t* colptr, const int64_t* rowval, const fdfd_c128* nzval, int base, SellHost& S, std::string& err) {
  S.n = n; S.nslices = (n + kSlice - 1) / kSlice;
  if (colptr[0] != base) { err = "colptr[0] does not equal the index base"; return FDFD_ERR_ARG; }
  const int64_t nnz = colptr[n] - base;
  if (nnz < 0) { err = "colptr is not non-decreasing"; return FDFD_ERR_ARG; }
  S.nnz = nnz;
  std::vector<int32_t> cnt(n, 0);
  for (int64_t j = 0; j < n; ++j) {
    if (colptr[j + 1] < colptr[j]) { err = "colptr is not non-decreasing"; return FDFD_ERR_ARG; }
    for (int64_t e = colptr[j] - base; e < colptr[j + 1] - base; ++e) {
      const int64_t i = rowval[e] - base;
      if (i < 0 || i >= n) { err = "rowval out of range"; return FDFD_ERR_ARG; }
      ++cnt[i];
    }
  }
  S.sptr.assign(S.nslices + 1, 0);
  for (int64_t s = 0; s < S.nslices; ++s) {
    int32_t w = 0;
    for (int64_t i = s * kSlice; i < std::min(n, (s + 1) * kSlice); ++i) w = std::max(w, cnt[i]);
    S.sptr[s + 1] = S.sptr[s] + (int64_t)w * kSlice;
  }
  const int64_t cap = S.sptr[S.nslices];
  S.col.resize(cap); S.val.assign(cap, c128(0.0, 0.0));
  for (int64_t s = 0; s < S.nslices; ++s)     // padding: gather the row's own x (always in range; rows past n gather x[n-1]) times 0
    for (int64_t e = S.sptr[s]; e < S.sptr[s + 1]; ++e) S.col[e] = (int32_t)std::min(n - 1, s * kSlice + (e - S.sptr[s]) % kSlice);
  std::vector<c128> diag(n, c128(0.0, 0.0));
  S.rowsum.assign(n, c128(0.0, 0.0));
  std::fill(cnt.begin(), cnt.end(), 0);
  for (int64_t j = 0; j < n; ++j)             // columns ascending => every row's entries end up sorted by column
    for (int64_t e = colptr[j] - base; e < colptr[j + 1] - base; ++e) {
      const int64_t i = rowval[e] - base;
      const c128 v(nzval[e].re, nzval[e].im);
      const int64_t s = i / kSlice, r = i % kSlice;
      const int64_t at = S.sptr[s] + (int64_t)cnt[i] * kSlice + r;
      S.col[at] = (int32_t)j; S.val[at] = v;
      ++cnt[i];
      if (i == j) diag[i] += v;
      S.rowsum[i] += v;
    }
  S.dinv.resize(n);
  for (int64_t i = 0; i < n; ++i) S.dinv[i] = norm2(diag[i]) > 0.0 ? crecip(diag[i]) : c128(1.0, 0.0);
  return FDFD_OK;
}

// one row of y = A x in the SELL layout (shared by the kernel and the host test hook: same order of summation)
__host__ __device__ __forceinline__ c128 sell_row(int64_t lo, int64_t hi, int r, const int32_t* __restrict__ col, const c128* __restrict__ val,
                                                  const c128* __restrict__ x) {
  c128 acc(0.0, 0.0);
  for (int64_t e = lo + r; e < hi; e += kSlice) cfma(acc, val[e], x[col[e]]);
  return acc;
}

// y = A x, one lane per row, one warp per slice (grid-stride over slices); fused dots as in k_apply:
//   NDOT 1: partials[block] = <d0, y>;  NDOT 2: partials[block] = (<y, d0>, |y|^2, 0)
template <int NDOT>
__global__ void __launch_bounds__(kSpThreads) k_sell_spmv(int64_t n, int64_t nslices, const int64_t* __restrict__ sptr, const int32_t* __restrict__ col,
                                                          const c128* __restrict__ val, const c128* __restrict__ x, c128* __restrict__ y,
                                                          const c128* __restrict__ d0, c128* __restrict__ partials, const int* __restrict__ done) {
  if (done && *done) return;
  const int lane = threadIdx.x & 31;
  const int64_t wpb = kSpThreads / 32;
  double acc[NDOT > 0 ? 2 * NDOT : 1];
#pragma unroll
  for (int k = 0; k < (NDOT > 0 ? 2 * NDOT : 1); ++k) acc[k] = 0.0;
  for (int64_t s = blockIdx.x * wpb + (threadIdx.x >> 5); s < nslices; s += (int64_t)gridDim.x * wpb) {
    const int64_t i = s * kSlice + lane;
    const c128 out = sell_row(sptr[s], sptr[s + 1], lane, col, val, x);
    if (i < n) {
      y[i] = out;
      if constexpr (NDOT == 1) {
        const c128 p = cmulc(d0[i], out);
        acc[0] += p.x; acc[1] += p.y;
      } else if constexpr (NDOT == 2) {
        const c128 p = cmulc(out, d0[i]);
        acc[0] += p.x; acc[1] += p.y;
        acc[2] += norm2(out);
      }
    }
  }
  if constexpr (NDOT > 0) block_reduce_store<kSpThreads, 2 * NDOT>(acc, reinterpret_cast<double*>(partials) + (size_t)blockIdx.x * 2 * NDOT);
}

// out = dinv .* in   (skipped once the solver's convergence flag is up, like every kernel of a captured iteration)
__global__ void k_diag_scale(int64_t n, const c128* __restrict__ dinv, const c128* __restrict__ in, c128* __restrict__ out, const int* __restrict__ done) {
  if (done && *done) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = dinv[i] * in[i];
}

// deterministic test vector with O(1) entries of both signs in both parts (operator comparison of the grid-hinted path)
__global__ void k_probe_fill(int64_t n, c128* __restrict__ v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    v[i] = c128((double)(h & 0xffffff) / 8388608.0 - 1.0, (double)((h >> 24) & 0xffffff) / 8388608.0 - 1.0);
  }
}
// partials[block] = (|a - b|^2, |a|^2)
__global__ void __launch_bounds__(kSpThreads) k_diff_norms(int64_t n, const c128* __restrict__ a, const c128* __restrict__ b, double* __restrict__ partials) {
  double acc[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)kSpThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSpThreads) {
    const c128 ai = a[i];
    acc[0] += norm2(ai - b[i]); acc[1] += norm2(ai);
  }
  block_reduce_store<kSpThreads, 2>(acc, partials + (size_t)blockIdx.x * 2);
}

struct SellDev {
  int64_t n = 0, nslices = 0;
  int blocks = 0;
  DevBuf<int64_t> sptr;
  DevBuf<int32_t> col;
  DevBuf<c128> val, dinv;
};

int sell_upload(fdfd_ctx* ctx, const SellHost& H, SellDev& D) {
  D.n = H.n; D.nslices = H.nslices;
  const int64_t wpb = kSpThreads / 32;
  D.blocks = (int)std::max<int64_t>(1, std::min<int64_t>((H.nslices + wpb - 1) / wpb, (int64_t)ctx->num_sms * 8));
  CUDA_TRY(ctx, D.sptr.alloc(H.sptr.size())); CUDA_TRY(ctx, D.col.alloc(std::max<size_t>(1, H.col.size())));
  CUDA_TRY(ctx, D.val.alloc(std::max<size_t>(1, H.val.size()))); CUDA_TRY(ctx, D.dinv.alloc(H.n));
  CUDA_TRY(ctx, cudaMemcpyAsync(D.sptr.p, H.sptr.data(), H.sptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  if (!H.col.empty()) {
    CUDA_TRY(ctx, cudaMemcpyAsync(D.col.p, H.col.data(), H.col.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(D.val.p, H.val.data(), H.val.size() * sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(D.dinv.p, H.dinv.data(), H.n * sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // H may go out of scope
  return FDFD_OK;
}

int sell_apply(fdfd_ctx* ctx, const SellDev& D, const c128* x, c128* y, const DotSpec& ds) {
  cudaStream_t st = ctx->stream;
#define SPMV(ND) k_sell_spmv<ND><<<D.blocks, kSpThreads, 0, st>>>(D.n, D.nslices, D.sptr.p, D.col.p, D.val.p, x, y, ds.d0, ds.partials, ds.done)
  switch (ds.ndot) {
    case 0: SPMV(0); break;
    case 1: SPMV(1); break;
    case 2: SPMV(2); break;
    default: fdfd_set_error(ctx, "sell_apply: ndot must be 0, 1 or 2"); return FDFD_ERR_ARG;
  }
#undef SPMV
  KLAUNCH(ctx);
  if (ds.nblocks_out) *ds.nblocks_out = D.blocks;
  return FDFD_OK;
}

}  // namespace

namespace {

using clk = std::chrono::steady_clock;
double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

// BiCGSTAB + Jacobi on the SELL image (the path of a matrix without a grid)
int solve_generic(fdfd_ctx* ctx, const SellDev& D, const fdfd_c128* b, fdfd_solve_opts_t o, fdfd_c128* x, fdfd_info_t* info, clk::time_point t0) {
  const int64_t n = D.n;
  o.solver = FDFD_SOLVER_BICGSTAB; o.precond = FDFD_PRECOND_JACOBI;
  KrylovWork W;
  FDFD_TRY(W.alloc(ctx, n, D.blocks, o.maxit, true));
  FDFD_TRY(fdfd_copy_in(ctx, W.b.p, b, (size_t)n * sizeof(c128)));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const double setup_ms = ms_since(t0);
  KrylovOps k;
  k.prec_f32 = false; k.prec_rhs = nullptr; k.fscale = 1.0;
  k.nab = D.blocks;
  k.apply = [ctx, &D](const void* xin, bool x_f32, c128* y, const DotSpec& ds) -> int {
    if (x_f32) { fdfd_set_error(ctx, "fdfd_dolinearsolve_csc: internal: fp32 input to the fp64 SpMV"); return FDFD_ERR_ARG; }
    return sell_apply(ctx, D, (const c128*)xin, y, ds);
  };
  k.precond = [ctx, &D, &W](bool hold, const void** out) -> int {
    c128* dst = hold ? W.ph.p : W.sh.p;
    *out = dst;
    k_diag_scale<<<W.nvec_blocks, 256, 0, ctx->stream>>>(D.n, D.dinv.p, hold ? W.p.p : W.s.p, dst, &W.scal.p->done); KLAUNCH(ctx);
    return FDFD_OK;
  };
  fdfd_info_t inf{};
  FDFD_TRY(krylov_bicgstab(ctx, W, k, o, &inf));
  inf.setup_ms = setup_ms;
  inf.mg_levels = 0;
  FDFD_TRY(fdfd_copy_out(ctx, x, W.x.p, (size_t)n * sizeof(c128)));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  inf.total_ms = ms_since(t0);
  if (info) *info = inf;
  if (inf.flag != FDFD_OK) {
    fdfd_set_error(ctx, "fdfd_dolinearsolve_csc: Krylov solver stopped with flag %d after %d iterations, relres %.3e", inf.flag, inf.iters, inf.relres);
    return inf.flag;
  }
  return FDFD_OK;
}

// (|a - b|^2, |a|^2) summed on the host from the per-CTA partials (fixed order)
int diff_norms(fdfd_ctx* ctx, int64_t n, const c128* a, const c128* b, DevBuf<double>& parts, int nb, double out[2]) {
  k_diff_norms<<<nb, kSpThreads, 0, ctx->stream>>>(n, a, b, parts.p); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  std::vector<double> h((size_t)nb * 2);
  CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), parts.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  out[0] = out[1] = 0.0;
  for (int i = 0; i < nb; ++i) { out[0] += h[2 * i]; out[1] += h[2 * i + 1]; }
  return FDFD_OK;
}

int check_csc_args(fdfd_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval, const fdfd_c128* nzval, int index_base,
                   const fdfd_c128* b, fdfd_c128* x) {
  ARG_CHECK(ctx, n >= 1 && n < ((int64_t)1 << 31), "n must be in [1, 2^31)");
  ARG_CHECK(ctx, colptr && rowval && nzval && b && x, "NULL argument");
  ARG_CHECK(ctx, index_base == 0 || index_base == 1, "index_base must be 0 or 1");
  return FDFD_OK;
}

}  // namespace

// dolinearsolve(A::SparseMatrixCSC{ComplexF64,Int64}, b, matrixsym) -> x   (src/solver/solver.jl:4-41; matrixsym is ignored there too, :29).
// colptr / rowval / nzval are HOST arrays exactly as Julia stores them (A.colptr, A.rowval, A.nzval; index_base 1) or 0-based;
// b and x may be host or device pointers.  Solver: BiCGSTAB + Jacobi on the SELL-32 image of A; opts->tol / maxit / check_every /
// use_graph / verbose are honoured, the preconditioner choice is not (a matrix has no grid to build a multigrid hierarchy from).
extern "C" int fdfd_dolinearsolve_csc(fdfd_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval, const fdfd_c128* nzval,
                                      int index_base, const fdfd_c128* b, const fdfd_solve_opts_t* opts, fdfd_c128* x, fdfd_info_t* info) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  FDFD_TRY(check_csc_args(ctx, n, colptr, rowval, nzval, index_base, b, x));
  const auto t0 = clk::now();
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  fdfd_solve_opts_t o;
  if (opts) o = *opts; else fdfd_default_opts(&o);
  SellDev D;
  {
    SellHost H;
    std::string err;
    const int st = csc_to_sell(n, colptr, rowval, nzval, index_base, H, err);
    if (st != FDFD_OK) { fdfd_set_error(ctx, "fdfd_dolinearsolve_csc: %s", err.c_str()); return st; }
    FDFD_TRY(sell_upload(ctx, H, D));
  }
  return solve_generic(ctx, D, b, o, x, info, t0);
}

// The same seam with the grid the matrix was assembled on (the caller of dolinearsolve always has one: solver.jl's callers build A
// from d.grid, nonlinear.jl:58-66).  If A turns out to BE the TM operator of that grid at omega for SOME permittivity -- the first
// solve of nonlinear.jl:66-69 and every Born step A + Diagonal(coeff |ez|^2) (nonlinear.jl:97) are -- the solve runs on the fast
// path: eps_eff = (A 1) / (w^2 eps0 L0) is read off the row sums (a difference stencil annihilates constants), the matrix-free
// operator + multigrid hierarchy are built from (grid, omega, eps_eff), and A v == A_matrixfree v is CHECKED on the device with a
// pseudo-random v (relative 1e-9, both derivative orderings tried) before the multigrid-preconditioned solver is trusted with it.
// Anything else (the 2N x 2N Gauss-Newton Jacobian of nonlinear.jl:120, a TE matrix, another grid) falls back to the generic path.
// The residual reported in info->relres is recomputed against the CALLER'S matrix.  info->mg_levels > 0 tells the fast path ran.
extern "C" int fdfd_dolinearsolve_csc_grid(fdfd_ctx* ctx, const fdfd_grid_t* g, double omega, int64_t n, const int64_t* colptr,
                                           const int64_t* rowval, const fdfd_c128* nzval, int index_base, const fdfd_c128* b,
                                           const fdfd_solve_opts_t* opts, fdfd_c128* x, fdfd_info_t* info) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  FDFD_TRY(check_csc_args(ctx, n, colptr, rowval, nzval, index_base, b, x));
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, omega > 0, "omega must be > 0");
  const auto t0 = clk::now();
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  fdfd_solve_opts_t o;
  if (opts) o = *opts; else fdfd_default_opts(&o);
  SellDev D;
  std::vector<c128> eps_eff;
  {
    SellHost H;
    std::string err;
    const int st = csc_to_sell(n, colptr, rowval, nzval, index_base, H, err);
    if (st != FDFD_OK) { fdfd_set_error(ctx, "fdfd_dolinearsolve_csc_grid: %s", err.c_str()); return st; }
    FDFD_TRY(sell_upload(ctx, H, D));
    if (n == g->Nx * g->Ny) {
      const double s = 1.0 / (omega * omega * kEps0 * g->L0);   // TM mass term: w^2 eps0 L0 eps_r (driven.jl:35, device.jl:40)
      eps_eff.resize(n);
      for (int64_t i = 0; i < n; ++i) eps_eff[i] = c128(s * H.rowsum[i].x, s * H.rowsum[i].y);
    }
  }
  const bool want_mg = (o.solver == FDFD_SOLVER_BICGSTAB || o.solver == FDFD_SOLVER_AUTO || o.solver == FDFD_SOLVER_MLKRYLOV) && o.precond == FDFD_PRECOND_MG;
  if (!eps_eff.empty() && want_mg) {
    for (int ordering : {FDFD_ORDER_FB, FDFD_ORDER_BF}) {
      fdfd_problem* P = nullptr;
      FDFD_TRY(fdfd_problem_create(ctx, g, FDFD_TM, ordering, omega, reinterpret_cast<const fdfd_c128*>(eps_eff.data()), &o, &P));
      struct Guard { fdfd_problem* p; ~Guard() { fdfd_problem_destroy(p); } } guard{P};
      // is A the matrix-free operator?  p = probe vector, v = A_matrixfree p, s = A p  (p, v, s are re-initialised by the solve)
      DevBuf<double> parts;
      const int nb = P->w.nvec_blocks;
      CUDA_TRY(ctx, parts.alloc((size_t)nb * 2));
      k_probe_fill<<<nb, 256, 0, ctx->stream>>>(n, P->w.p.p); KLAUNCH(ctx);
      DotSpec d0;
      FDFD_TRY(launch_apply(ctx, P->op.view(), false, P->w.p.p, false, P->w.v.p, d0));
      FDFD_TRY(sell_apply(ctx, D, P->w.p.p, P->w.s.p, d0));
      double nn[2];
      FDFD_TRY(diff_norms(ctx, n, P->w.s.p, P->w.v.p, parts, nb, nn));
      const double opdiff = nn[1] > 0.0 ? std::sqrt(nn[0] / nn[1]) : 1.0;
      if (o.verbose) fprintf(stderr, "[fdfd_b200] dolinearsolve: |A v - A_matrixfree v| / |A v| = %.3e (ordering %d)\n", opdiff, ordering);
      if (!(opdiff <= 1e-9)) continue;
      FDFD_TRY(fdfd_problem_set_rhs(P, b));
      fdfd_info_t inf{};
      FDFD_TRY(fdfd_problem_solve(P, &inf));
      // residual against the caller's matrix: t = A x
      FDFD_TRY(sell_apply(ctx, D, P->w.x.p, P->w.t.p, d0));
      FDFD_TRY(diff_norms(ctx, n, P->w.b.p, P->w.t.p, parts, nb, nn));
      inf.relres = nn[1] > 0.0 ? std::sqrt(nn[0] / nn[1]) : 0.0;
      if (inf.flag == FDFD_OK && !(inf.relres <= 10.0 * o.tol)) inf.flag = FDFD_ERR_NOCONV;
      FDFD_TRY(fdfd_problem_get_solution(P, x));
      inf.total_ms = ms_since(t0);
      if (info) *info = inf;
      if (inf.flag != FDFD_OK) {
        fdfd_set_error(ctx, "fdfd_dolinearsolve_csc_grid: solver stopped with flag %d after %d iterations, relres %.3e", inf.flag, inf.iters, inf.relres);
        return inf.flag;
      }
      return FDFD_OK;
    }
  }
  return solve_generic(ctx, D, b, o, x, info, t0);
}

// measurement hook (GPU): average duration in ms of `reps` launches of the SELL-32 SpMV with the two fused dots (the variant the
// BiCGSTAB loop runs most), CUDA events on the ctx stream after 3 warm-up launches; *alg_bytes = algorithmic bytes per launch
// (per stored entry 16 + 4 + a 16-byte gathered x; per row 16 for y + 16 for the dot operand)
extern "C" int fdfd_debug_sell_bench(fdfd_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval, const fdfd_c128* nzval,
                                     int index_base, int reps, double* ms_per_launch, double* alg_bytes) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  ARG_CHECK(ctx, n >= 1 && n < ((int64_t)1 << 31) && colptr && rowval && nzval && reps >= 1 && ms_per_launch, "bad arguments");
  ARG_CHECK(ctx, index_base == 0 || index_base == 1, "index_base must be 0 or 1");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  SellDev D;
  int64_t stored = 0;
  {
    SellHost H;
    std::string err;
    const int st = csc_to_sell(n, colptr, rowval, nzval, index_base, H, err);
    if (st != FDFD_OK) { fdfd_set_error(ctx, "fdfd_debug_sell_bench: %s", err.c_str()); return st; }
    stored = H.sptr[H.nslices];
    FDFD_TRY(sell_upload(ctx, H, D));
  }
  DevBuf<c128> x, y, d, parts;
  CUDA_TRY(ctx, x.alloc(n)); CUDA_TRY(ctx, y.alloc(n)); CUDA_TRY(ctx, d.alloc(n)); CUDA_TRY(ctx, parts.alloc((size_t)D.blocks * 2));
  const int nb = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 8);
  k_probe_fill<<<nb, 256, 0, ctx->stream>>>(n, x.p); KLAUNCH(ctx);
  k_probe_fill<<<nb, 256, 0, ctx->stream>>>(n, d.p); KLAUNCH(ctx);
  DotSpec ds; ds.ndot = 2; ds.d0 = d.p; ds.partials = parts.p;
  for (int i = 0; i < 3; ++i) FDFD_TRY(sell_apply(ctx, D, x.p, y.p, ds));
  cudaEvent_t e0, e1;
  CUDA_TRY(ctx, cudaEventCreate(&e0)); CUDA_TRY(ctx, cudaEventCreate(&e1));
  CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
  int st = FDFD_OK;
  for (int i = 0; i < reps && st == FDFD_OK; ++i) st = sell_apply(ctx, D, x.p, y.p, ds);
  cudaEventRecord(e1, ctx->stream);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  FDFD_TRY(st);
  CUDA_TRY(ctx, cudaGetLastError());
  *ms_per_launch = (double)ms / reps;
  if (alg_bytes) *alg_bytes = 36.0 * (double)stored + 32.0 * (double)n;
  return FDFD_OK;
}

// host-only test hook (no GPU needed): y = A x through the SAME CSC -> SELL-32 transposition and the same per-row summation as the
// kernel; also returns the padded entry count and the inverse diagonal the Jacobi preconditioner would use
extern "C" int fdfd_debug_sell_spmv(int64_t n, const int64_t* colptr, const int64_t* rowval, const fdfd_c128* nzval, int index_base,
                                    const fdfd_c128* x, fdfd_c128* y, fdfd_c128* dinv, fdfd_c128* rowsum, int64_t* padded_entries) {
  if (n < 1 || n >= ((int64_t)1 << 31) || !colptr || !rowval || !nzval || !x || !y || !(index_base == 0 || index_base == 1)) return FDFD_ERR_ARG;
  SellHost H;
  std::string err;
  const int st = csc_to_sell(n, colptr, rowval, nzval, index_base, H, err);
  if (st != FDFD_OK) return st;
  const c128* xx = reinterpret_cast<const c128*>(x);
  c128* yy = reinterpret_cast<c128*>(y);
  for (int64_t s = 0; s < H.nslices; ++s)
    for (int r = 0; r < kSlice; ++r) {
      const int64_t i = s * kSlice + r;
      const c128 out = sell_row(H.sptr[s], H.sptr[s + 1], r, H.col.data(), H.val.data(), xx);
      if (i < n) yy[i] = out;
    }
  if (dinv) std::memcpy(dinv, H.dinv.data(), sizeof(c128) * n);
  if (rowsum) std::memcpy(rowsum, H.rowsum.data(), sizeof(c128) * n);
  if (padded_entries) *padded_entries = H.sptr[H.nslices];
  return FDFD_OK;
}
