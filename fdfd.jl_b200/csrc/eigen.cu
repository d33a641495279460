// eigen.cu -- eigenfrequency(d, pol, nev; which) (src/solver/eigen.jl:69-115) as shift-invert Arnoldi.
// The reference hands A and sigma to Arpack.eigs, which factorises (A - sigma I) once (UMFPACK) and runs
// implicitly restarted Arnoldi on OP = (A - sigma I)^-1.  Here OP is applied by the same GPU Krylov solve
// as the driven problem, because
//   TM:  (Teps_r^-1 L - sigma) y = v  <=>  (L/mu0 + w0^2 eps0 eps_r) y = eps_r v / mu0   == driven TM operator at w0
//   TE:  (A - sigma) y = v,  sigma = -w0^2 mu0                                         == driven TE operator at w0
// Orthogonalisation (classical Gram-Schmidt, applied twice) runs on the device against a basis resident in
// HBM; only the small projected matrix lives on the host (arnoldi.cu: Krylov-Schur thick restarts keep the basis at
// ncv + 1 vectors like Arpack's implicit restarts; a complex shifted-QR iteration gives the Ritz pairs).
// `which` is interpreted on the transformed spectrum nu = 1/(lambda - sigma) like ARPACK.
#include "arnoldi.cuh"
#include "krylov.cuh"
#include "reduce.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>

using cd = std::complex<double>;

namespace {

constexpr int kT = 256;

__global__ void __launch_bounds__(kT) k_dotc(int64_t N, const c128* __restrict__ a, const c128* __restrict__ b, double* __restrict__ partials) {
  double acc[2] = {0, 0};
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < N; i += (int64_t)gridDim.x * kT) {
    const c128 q = cmulc(a[i], b[i]);
    acc[0] += q.x; acc[1] += q.y;
  }
  block_reduce_store<kT, 2>(acc, partials + (size_t)blockIdx.x * 2);
}
__global__ void k_dot_final(const double* __restrict__ partials, int nb, c128* __restrict__ out, int accumulate) {
  double res[2];
  final_reduce<kT, 2>(partials, nb, res);
  if (threadIdx.x == 0) { if (accumulate) { out->x += res[0]; out->y += res[1]; } else *out = c128(res[0], res[1]); }
}
// out (+)= c * v
__global__ void k_axpy_host(int64_t N, c128 c, const c128* __restrict__ v, c128* __restrict__ out, int first) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = first ? c * v[i] : out[i] + c * v[i];
}
// inner right-hand side: TM  b = eps_r .* v / mu0 ; TE  b = v
__global__ void k_eig_rhs(int64_t N, int tm, double inv_mu0, const c128* __restrict__ eps, const c128* __restrict__ v, c128* __restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    if (tm) { const c128 q = eps[i] * v[i]; b[i] = c128(q.x * inv_mu0, q.y * inv_mu0); }
    else b[i] = v[i];
  }
}
__global__ void k_seed(int64_t N, c128* __restrict__ x, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    x[i] = c128((double)(z & 0xFFFFFFFF) / 4294967296.0 - 0.5, (double)(z >> 32) / 4294967296.0 - 0.5);
  }
}

}  // namespace

// ---- host: eigen-decomposition of a small complex upper-Hessenberg matrix (column major, ld = n) -----------
// (external linkage: shared with slab_multi.cu)
// shifted QR with Givens rotations -> Schur form T = Z^H H Z, then eigenvectors of T by back substitution.
bool hess_eig(int n, std::vector<cd> H, std::vector<cd>& evals, std::vector<cd>& evecs) {
  auto at = [&](std::vector<cd>& M, int i, int j) -> cd& { return M[(size_t)j * n + i]; };
  std::vector<cd> Z((size_t)n * n, cd(0, 0));
  for (int i = 0; i < n; ++i) at(Z, i, i) = 1.0;
  const double eps = 2.2e-16;
  int ihi = n - 1, iter = 0, total = 0;
  std::vector<cd> cs(n), sn(n);
  while (ihi >= 0) {
    int l = ihi;
    while (l > 0) {
      const double sc = std::abs(at(H, l - 1, l - 1)) + std::abs(at(H, l, l));
      if (std::abs(at(H, l, l - 1)) <= eps * (sc > 0 ? sc : 1.0)) { at(H, l, l - 1) = 0.0; break; }
      --l;
    }
    if (l == ihi) { --ihi; iter = 0; continue; }
    if (++total > 60 * n) return false;
    ++iter;
    cd mu;
    if (iter % 11 == 10) mu = at(H, ihi, ihi) + std::abs(at(H, ihi, ihi - 1));  // exceptional shift
    else {  // Wilkinson: eigenvalue of the trailing 2x2 closer to H[ihi,ihi]
      const cd a = at(H, ihi - 1, ihi - 1), b = at(H, ihi - 1, ihi), c = at(H, ihi, ihi - 1), d = at(H, ihi, ihi);
      const cd tr = a + d, det = a * d - b * c;
      const cd disc = std::sqrt(tr * tr - 4.0 * det);
      const cd e1 = 0.5 * (tr + disc), e2 = 0.5 * (tr - disc);
      mu = std::abs(e1 - d) < std::abs(e2 - d) ? e1 : e2;
    }
    for (int k = l; k <= ihi; ++k) at(H, k, k) -= mu;
    for (int k = l; k < ihi; ++k) {
      const cd a = at(H, k, k), b = at(H, k + 1, k);
      const double r = std::hypot(std::abs(a), std::abs(b));
      cd c = 1.0, s = 0.0;
      if (r > 0) { c = a / r; s = b / r; }
      cs[k] = c; sn[k] = s;
      for (int j = k; j < n; ++j) {
        const cd x = at(H, k, j), y = at(H, k + 1, j);
        at(H, k, j) = std::conj(c) * x + std::conj(s) * y;
        at(H, k + 1, j) = -s * x + c * y;
      }
    }
    for (int k = l; k < ihi; ++k) {
      const cd c = cs[k], s = sn[k];
      const int top = std::min(k + 2, ihi);
      for (int i = 0; i <= top; ++i) {
        const cd x = at(H, i, k), y = at(H, i, k + 1);
        at(H, i, k) = x * c + y * s;
        at(H, i, k + 1) = -x * std::conj(s) + y * std::conj(c);
      }
      for (int i = 0; i < n; ++i) {
        const cd x = at(Z, i, k), y = at(Z, i, k + 1);
        at(Z, i, k) = x * c + y * s;
        at(Z, i, k + 1) = -x * std::conj(s) + y * std::conj(c);
      }
    }
    for (int k = l; k <= ihi; ++k) at(H, k, k) += mu;
  }
  evals.resize(n);
  evecs.assign((size_t)n * n, cd(0, 0));
  double tnorm = 0;
  for (int i = 0; i < n; ++i) { evals[i] = at(H, i, i); tnorm = std::max(tnorm, std::abs(evals[i])); }
  std::vector<cd> y(n);
  for (int k = 0; k < n; ++k) {
    std::fill(y.begin(), y.end(), cd(0, 0));
    y[k] = 1.0;
    for (int i = k - 1; i >= 0; --i) {
      cd sacc = 0;
      for (int j = i + 1; j <= k; ++j) sacc += at(H, i, j) * y[j];
      cd den = at(H, i, i) - at(H, k, k);
      if (std::abs(den) < eps * tnorm) den = eps * tnorm;
      y[i] = -sacc / den;
    }
    double nrm = 0;
    std::vector<cd> x(n, cd(0, 0));
    for (int j = 0; j <= k; ++j) for (int i = 0; i < n; ++i) x[i] += at(Z, i, j) * y[j];
    for (int i = 0; i < n; ++i) nrm += std::norm(x[i]);
    nrm = std::sqrt(nrm);
    for (int i = 0; i < n; ++i) evecs[(size_t)k * n + i] = x[i] / nrm;
  }
  return true;
}

double which_key(int which, cd nu) {
  switch (which) {
    case FDFD_WHICH_LR: return nu.real();
    case FDFD_WHICH_SR: return -nu.real();
    case FDFD_WHICH_LI: return nu.imag();
    case FDFD_WHICH_SI: return -nu.imag();
    default: return std::abs(nu);
  }
}

extern "C" int fdfd_eigenfrequency(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, double omega0, int nev, int which, int ncv,
                                   const fdfd_c128* eps_r, const fdfd_solve_opts_t* opts, fdfd_c128* omega_out,
                                   fdfd_c128* fields, fdfd_info_t* info) {
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, pol == FDFD_TM || pol == FDFD_TE, "pol must be FDFD_TM or FDFD_TE");
  ARG_CHECK(ctx, nev >= 1 && eps_r && omega_out, "bad arguments");
  ARG_CHECK(ctx, which >= FDFD_WHICH_LM && which <= FDFD_WHICH_SI, "bad `which`");
  ARG_CHECK(ctx, omega0 > 0, "omega0 must be > 0");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int64_t N = g->Nx * g->Ny;
  ARG_CHECK(ctx, nev + 2 <= N, "nev too large for the grid");
  fdfd_solve_opts_t o;
  if (opts) o = *opts; else fdfd_default_opts(&o);
  // Arpack's default tolerance (machine epsilon, eigen.jl:86 passes none) presumes an exact factorisation behind OP; here OP is
  // an iterative solve, so the Ritz residual can only be asked to reach a small multiple of the inner tolerance.  The inner
  // solves run to min(opts.tol, 1e-11) and a pair counts as converged at |b^T y| <= 10 x that x |nu|: 1e-10 by default (the
  // eigenfrequencies then agree with the oracle's Arpack run to ~1e-10, bar 1e-8), tighter if the caller passes a tighter opts.tol.
  o.tol = std::min(o.tol, 1e-11);
  const double tol_eig = 10.0 * o.tol;
  if (ncv <= 0) ncv = std::max(20, 2 * nev + 1);  // Arpack.jl default
  ncv = (int)std::min<int64_t>(std::max(ncv, nev + 2), N - 1);
  const int max_steps = std::max(300, 30 * nev) + ncv;   // Arpack.jl: maxiter = 300 restarts; here a bound on operator applications
  const double eps0 = kEps0 * g->L0, mu0 = kMu0 * g->L0;
  const cd sigma = pol == FDFD_TM ? cd(-omega0 * omega0 * mu0 * eps0, 0) : cd(-omega0 * omega0 * mu0, 0);  // eigen.jl:86,104

  fdfd_problem* P = nullptr;
  FDFD_TRY(fdfd_problem_create(ctx, g, pol, FDFD_ORDER_FB, omega0, eps_r, &o, &P));
  struct Guard { fdfd_problem* p; ~Guard() { fdfd_problem_destroy(p); } } guard{P};
  cudaStream_t st = ctx->stream;
  const int nb = P->w.nvec_blocks;
  std::vector<DevBuf<c128>> V;  // Arnoldi basis: never more than ncv + 1 vectors (+ the k rotated ones during a restart)
  DevBuf<c128> wv, scal1; DevBuf<double> parts;
  CUDA_TRY(ctx, wv.alloc(N)); CUDA_TRY(ctx, scal1.alloc(1)); CUDA_TRY(ctx, parts.alloc((size_t)nb * 2));
  int inner_its = 0;
  const int64_t launches0 = ctx->launches;
  double inner_ms = 0;

  auto ensure = [&](int j) -> int {
    while ((int)V.size() <= j) {
      V.emplace_back();
      if (V.back().alloc(N) != cudaSuccess) { cudaGetLastError(); fdfd_set_error(ctx, "out of device memory for the Arnoldi basis (%d vectors of %lld points)", j + 1, (long long)N); return FDFD_ERR_ALLOC; }
    }
    return FDFD_OK;
  };
  ArnoldiOps ops;
  // w = OP v_j : inner Krylov solve with the driven operator
  ops.op_apply = [&](int j) -> int {
    k_eig_rhs<<<nb, 256, 0, st>>>(N, pol == FDFD_TM, 1.0 / mu0, P->op.eps.p, V[j].p, P->w.b.p); KLAUNCH(ctx);
    P->have_rhs = true;
    fdfd_info_t inf{};
    FDFD_TRY(fdfd_problem_solve(P, &inf));
    if (inf.flag != FDFD_OK) { fdfd_set_error(ctx, "fdfd_eigenfrequency: inner solve %d failed (flag %d, relres %.2e)", j, inf.flag, inf.relres); return inf.flag; }
    inner_its += inf.iters; inner_ms += inf.solve_ms;
    CUDA_TRY(ctx, cudaMemcpyAsync(wv.p, P->w.x.p, N * sizeof(c128), cudaMemcpyDeviceToDevice, st));
    return FDFD_OK;
  };
  ops.dot_v_w = [&](int i, cd* h) -> int {
    k_dotc<<<nb, kT, 0, st>>>(N, V[i].p, wv.p, parts.p); KLAUNCH(ctx);
    k_dot_final<<<1, kT, 0, st>>>(parts.p, nb, scal1.p, 0); KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(h, scal1.p, sizeof(c128), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return FDFD_OK;
  };
  ops.axpy_w = [&](int i, cd h) -> int {
    k_axpy_host<<<nb, 256, 0, st>>>(N, to_c128(-h), V[i].p, wv.p, 0); KLAUNCH(ctx);
    return FDFD_OK;
  };
  ops.norm_w = [&](double* nrm) -> int {
    cd h;
    k_dotc<<<nb, kT, 0, st>>>(N, wv.p, wv.p, parts.p); KLAUNCH(ctx);
    k_dot_final<<<1, kT, 0, st>>>(parts.p, nb, scal1.p, 0); KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(&h, scal1.p, sizeof(c128), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    *nrm = std::sqrt(std::max(0.0, h.real()));
    return FDFD_OK;
  };
  ops.set_v = [&](int j, double s) -> int {
    FDFD_TRY(ensure(j));
    k_axpy_host<<<nb, 256, 0, st>>>(N, c128(s, 0.0), wv.p, V[j].p, 1); KLAUNCH(ctx);
    return FDFD_OK;
  };
  ops.random_w = [&](int seed) -> int {
    k_seed<<<nb, 256, 0, st>>>(N, wv.p, 20260101ull + 7919ull * (uint64_t)seed); KLAUNCH(ctx);
    return FDFD_OK;
  };
  ops.rotate_basis = [&](int m, int k, const std::vector<cd>& Q) -> int {
    std::vector<DevBuf<c128>> T(k);
    for (int i = 0; i < k; ++i) {
      if (T[i].alloc(N) != cudaSuccess) { cudaGetLastError(); fdfd_set_error(ctx, "out of device memory for the thick restart"); return FDFD_ERR_ALLOC; }
      for (int j = 0; j < m; ++j) { k_axpy_host<<<nb, 256, 0, st>>>(N, to_c128(Q[(size_t)i * m + j]), V[j].p, T[i].p, j == 0); KLAUNCH(ctx); }
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    std::swap(V[k], V[m]);                         // v_{k+1} <- v_{m+1} (V[k], if k < m, is consumed above)
    for (int i = 0; i < k; ++i) std::swap(V[i], T[i]);
    return FDFD_OK;   // T (the old vectors) is released here
  };
  ArnoldiResult R;
  FDFD_TRY(krylov_schur(ctx, ops, nev, ncv, which, tol_eig, max_steps, o.verbose != 0, R));
  const int m = R.m;

  // eigenvalues: lambda = sigma + 1/nu ; TM ω = sqrt(-λ/μ₀/ϵ₀) (eigen.jl:87), TE ω = sqrt(-λ/μ₀) (eigen.jl:105)
  DevBuf<c128> ez, f3;
  CUDA_TRY(ctx, ez.alloc(N));
  if (fields) CUDA_TRY(ctx, f3.alloc(3 * N));
  for (int e = 0; e < nev; ++e) {
    const cd lam = sigma + 1.0 / R.nu[e];
    const cd om = pol == FDFD_TM ? std::sqrt(-lam / mu0 / eps0) : std::sqrt(-lam / mu0);
    omega_out[e].re = om.real(); omega_out[e].im = om.imag();
    if (!fields) continue;
    for (int i = 0; i < m; ++i) {
      k_axpy_host<<<nb, 256, 0, st>>>(N, to_c128(R.Y[(size_t)e * m + i]), V[i].p, ez.p, i == 0); KLAUNCH(ctx);
    }
    // TM: H from FORWARD stretched differences (eigen.jl:90-91); TE: backward, eps averaging swapped (eigen.jl:108-109)
    FDFD_TRY(launch_recover(ctx, P->op, ez.p, pol == FDFD_TM ? 1 : 0, om, 1, f3.p));
    FDFD_TRY(fdfd_copy_out(ctx, fields + (size_t)e * 3 * N, f3.p, 3 * N * sizeof(c128)));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
  }
  if (info) {
    std::memset(info, 0, sizeof(*info));
    info->iters = inner_its; info->flag = FDFD_OK; info->relres = o.tol; info->solve_ms = inner_ms;
    info->setup_ms = P->setup_ms; info->launches = ctx->launches - launches0; info->restarts = R.steps;  // restarts := Arnoldi steps (operator applications)
    info->mg_levels = P->mgf ? P->mgf->levels() : (P->mgd ? P->mgd->levels() : 0);
    info->total_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  }
  return FDFD_OK;
}

// host-only test hook: eigen-decomposition used for the Ritz pairs (column-major n x n Hessenberg in, eigenvalues and
// unit-norm eigenvectors out).  Runs without a GPU; lets the CPU test-suite validate the small dense solver.
extern "C" int fdfd_debug_hess_eig(int n, const fdfd_c128* H, fdfd_c128* evals, fdfd_c128* evecs) {
  if (n < 1 || !H || !evals || !evecs) return FDFD_ERR_ARG;
  std::vector<cd> h((size_t)n * n), ev, vec;
  std::memcpy(h.data(), H, sizeof(cd) * n * n);
  if (!hess_eig(n, h, ev, vec)) return FDFD_ERR_NOCONV;
  std::memcpy(evals, ev.data(), sizeof(cd) * n);
  std::memcpy(evecs, vec.data(), sizeof(cd) * n * n);
  return FDFD_OK;
}
