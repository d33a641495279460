// arnoldi.cu -- Krylov-Schur driver (see arnoldi.cuh).  New work: the reference delegates to Arpack.eigs
// (src/solver/eigen.jl:86,104); Arpack.jl is not part of /root/reference, what is restated here is its published
// contract: nev Ritz pairs of OP = (A - sigma I)^-1 selected by `which` on OP's spectrum, a basis bounded by ncv.
//
// One cycle: expand the Krylov decomposition  OP V_m = V_m S_m + v_{m+1} b^T  to m = ncv by Arnoldi steps (classical
// Gram-Schmidt applied twice), take the Ritz pairs of S_m, test |b^T y| <= tol |nu| for the nev wanted ones, otherwise keep the
// k = nev + (ncv - nev)/2 best Ritz vectors: with Q an orthonormal basis of their coefficient vectors,
//     V <- V Q,  S <- Q^H S Q (k x k),  b <- Q^T b,  v_{k+1} <- v_{m+1}
// is again a Krylov decomposition (span Q is an invariant subspace of S_m), and the expansion continues from column k.
// After a restart S is no longer Hessenberg (full k x k block + dense row k), hence general_eig.
#include "arnoldi.cuh"
#include "common.cuh"
#include <algorithm>
#include <cmath>

using cd = std::complex<double>;

bool general_eig(int n, std::vector<cd> A, std::vector<cd>& evals, std::vector<cd>& evecs) {
  auto at = [n](std::vector<cd>& M, int i, int j) -> cd& { return M[(size_t)j * n + i]; };
  std::vector<cd> P((size_t)n * n, cd(0, 0));
  for (int i = 0; i < n; ++i) at(P, i, i) = 1.0;
  std::vector<cd> v(n);
  for (int k = 0; k + 2 < n; ++k) {
    const int r = n - (k + 1);   // length of the reflector
    double xn = 0;
    for (int i = 0; i < r; ++i) xn += std::norm(at(A, k + 1 + i, k));
    double below = xn - std::norm(at(A, k + 1, k));
    xn = std::sqrt(xn);
    if (!(below > 0.0)) continue;   // column already in Hessenberg form
    const cd x0 = at(A, k + 1, k);
    const cd phase = std::abs(x0) > 0 ? x0 / std::abs(x0) : cd(1.0, 0.0);
    const cd alpha = -phase * xn;
    double vn = 0;
    for (int i = 0; i < r; ++i) { v[i] = at(A, k + 1 + i, k); if (i == 0) v[i] -= alpha; vn += std::norm(v[i]); }
    vn = std::sqrt(vn);
    if (!(vn > 0.0)) continue;
    for (int i = 0; i < r; ++i) v[i] /= vn;
    for (int j = 0; j < n; ++j) {   // A <- (I - 2 v v^H) A
      cd s = 0;
      for (int i = 0; i < r; ++i) s += std::conj(v[i]) * at(A, k + 1 + i, j);
      for (int i = 0; i < r; ++i) at(A, k + 1 + i, j) -= 2.0 * v[i] * s;
    }
    for (int i = 0; i < n; ++i) {   // A <- A (I - 2 v v^H),  P <- P (I - 2 v v^H)
      cd s = 0, sp = 0;
      for (int j = 0; j < r; ++j) { s += at(A, i, k + 1 + j) * v[j]; sp += at(P, i, k + 1 + j) * v[j]; }
      for (int j = 0; j < r; ++j) { at(A, i, k + 1 + j) -= 2.0 * s * std::conj(v[j]); at(P, i, k + 1 + j) -= 2.0 * sp * std::conj(v[j]); }
    }
    at(A, k + 1, k) = alpha;
    for (int i = k + 2; i < n; ++i) at(A, i, k) = 0.0;
  }
  std::vector<cd> hv;
  if (!hess_eig(n, A, evals, hv)) return false;
  evecs.assign((size_t)n * n, cd(0, 0));
  for (int e = 0; e < n; ++e) {
    double nrm = 0;
    for (int i = 0; i < n; ++i) {
      cd s = 0;
      for (int j = 0; j < n; ++j) s += at(P, i, j) * hv[(size_t)e * n + j];
      evecs[(size_t)e * n + i] = s;
      nrm += std::norm(s);
    }
    nrm = std::sqrt(nrm);
    if (nrm > 0) for (int i = 0; i < n; ++i) evecs[(size_t)e * n + i] /= nrm;
  }
  return true;
}

int krylov_schur(fdfd_ctx* ctx, const ArnoldiOps& ops, int nev, int ncv, int which, double tol, int max_steps, bool verbose,
                 ArnoldiResult& out) {
  ARG_CHECK(ctx, nev >= 1 && ncv >= nev + 2, "Krylov-Schur: need ncv >= nev + 2");
  const int ld = ncv + 1;
  std::vector<cd> S((size_t)ld * ncv, cd(0, 0));   // column major: S[j ld + i], rows 0..m (row m = b^T)
  int m = 0;
  out = ArnoldiResult();
  std::vector<cd> evals, evecs;
  std::vector<int> idx;
  std::vector<double> res;

  // Ritz pairs of S_m, residual estimates |b^T y|, `which` order; true when the first nev are converged
  auto ritz = [&](bool* ok) -> int {
    std::vector<cd> Sm((size_t)m * m);
    for (int j = 0; j < m; ++j) for (int i = 0; i < m; ++i) Sm[(size_t)j * m + i] = S[(size_t)j * ld + i];
    if (!general_eig(m, Sm, evals, evecs)) { fdfd_set_error(ctx, "Krylov-Schur: QR iteration on the projected matrix did not converge"); return FDFD_ERR_NOCONV; }
    res.assign(m, 0.0);
    for (int e = 0; e < m; ++e) {
      cd s = 0;
      for (int j = 0; j < m; ++j) s += S[(size_t)j * ld + m] * evecs[(size_t)e * m + j];
      res[e] = std::abs(s);
    }
    idx.resize(m);
    for (int i = 0; i < m; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return which_key(which, evals[a]) > which_key(which, evals[b]); });
    *ok = m >= nev;
    for (int e = 0; e < std::min(nev, m); ++e) if (!(res[idx[e]] <= tol * std::abs(evals[idx[e]]))) *ok = false;
    return FDFD_OK;
  };
  // w <- w - V (V^H w), twice; coefficients accumulated into col (may be null)
  auto orthogonalise = [&](int nvec, cd* col) -> int {
    for (int pass = 0; pass < 2; ++pass)
      for (int i = 0; i < nvec; ++i) {
        cd h;
        FDFD_TRY(ops.dot_v_w(i, &h));
        FDFD_TRY(ops.axpy_w(i, h));
        if (col) col[i] += h;
      }
    return FDFD_OK;
  };

  // v_1: normalised pseudo-random vector (ARPACK starts from a random residual vector)
  {
    FDFD_TRY(ops.random_w(0));
    double n0 = 0;
    FDFD_TRY(ops.norm_w(&n0));
    ARG_CHECK(ctx, n0 > 0, "Krylov-Schur: zero start vector");
    FDFD_TRY(ops.set_v(0, 1.0 / n0));
  }
  bool converged = false;
  while (!converged) {
    while (m < ncv) {
      if (out.steps >= max_steps) {
        fdfd_set_error(ctx, "Krylov-Schur: %d Ritz pairs did not converge within %d operator applications (%d restarts)", nev, out.steps, out.restarts);
        return FDFD_ERR_NOCONV;
      }
      FDFD_TRY(ops.op_apply(m));
      ++out.steps;
      cd* col = &S[(size_t)m * ld];
      FDFD_TRY(orthogonalise(m + 1, col));
      double hn = 0, cn = 0;
      FDFD_TRY(ops.norm_w(&hn));
      for (int i = 0; i <= m; ++i) cn += std::norm(col[i]);
      const bool breakdown = !(hn > 1e-14 * std::sqrt(cn));
      col[m + 1] = breakdown ? 0.0 : hn;
      ++m;
      if (breakdown) {
        // span V_m is invariant under OP: its Ritz pairs are exact.  If it holds fewer than nev of them, continue in a fresh
        // direction orthogonal to it (the decomposition stays valid with b = 0).
        if (verbose) fprintf(stderr, "[fdfd_b200] krylov-schur: invariant subspace of dimension %d\n", m);
        FDFD_TRY(ops.random_w(out.steps + 1));
        FDFD_TRY(orthogonalise(m, nullptr));
        double rn = 0;
        FDFD_TRY(ops.norm_w(&rn));
        ARG_CHECK(ctx, rn > 0, "Krylov-Schur: could not leave an invariant subspace");
        FDFD_TRY(ops.set_v(m, 1.0 / rn));
      } else {
        FDFD_TRY(ops.set_v(m, 1.0 / hn));
      }
      if (m >= nev + 2 || m == ncv || (breakdown && m >= nev)) {
        bool ok = false;
        FDFD_TRY(ritz(&ok));
        if (verbose) {
          double worst = 0;
          for (int e = 0; e < std::min(nev, m); ++e) worst = std::max(worst, res[idx[e]] / std::abs(evals[idx[e]]));
          fprintf(stderr, "[fdfd_b200] krylov-schur: m=%d steps=%d restarts=%d worst wanted |b^T y|/|nu| %.2e converged=%d\n", m, out.steps, out.restarts, worst, (int)ok);
        }
        if (ok) { converged = true; break; }
      }
    }
    if (converged) break;
    // ---- thick restart: keep the k best Ritz vectors (Ritz pairs of S_ncv were just computed)
    int k = std::min(m - 1, nev + std::max(1, (ncv - nev) / 2));
    std::vector<cd> Q((size_t)m * k);
    int kk = 0;
    for (int c = 0; c < k; ++c) {   // modified Gram-Schmidt (twice) of the wanted eigenvectors; a dependent one is dropped
      cd* q = &Q[(size_t)kk * m];
      for (int i = 0; i < m; ++i) q[i] = evecs[(size_t)idx[c] * m + i];
      for (int pass = 0; pass < 2; ++pass)
        for (int p = 0; p < kk; ++p) {
          const cd* qp = &Q[(size_t)p * m];
          cd s = 0;
          for (int i = 0; i < m; ++i) s += std::conj(qp[i]) * q[i];
          for (int i = 0; i < m; ++i) q[i] -= s * qp[i];
        }
      double nn = 0;
      for (int i = 0; i < m; ++i) nn += std::norm(q[i]);
      nn = std::sqrt(nn);
      if (nn < 1e-8) continue;
      for (int i = 0; i < m; ++i) q[i] /= nn;
      ++kk;
    }
    k = kk;
    ARG_CHECK(ctx, k >= 1, "Krylov-Schur: no Ritz vector left to restart with");
    Q.resize((size_t)m * k);
    std::vector<cd> SQ((size_t)m * k, cd(0, 0)), Sn((size_t)k * k, cd(0, 0)), bn(k, cd(0, 0));
    for (int c = 0; c < k; ++c)
      for (int j = 0; j < m; ++j) {
        const cd qjc = Q[(size_t)c * m + j];
        for (int i = 0; i < m; ++i) SQ[(size_t)c * m + i] += S[(size_t)j * ld + i] * qjc;
        bn[c] += S[(size_t)j * ld + m] * qjc;
      }
    for (int c = 0; c < k; ++c)
      for (int r = 0; r < k; ++r) {
        cd s = 0;
        for (int i = 0; i < m; ++i) s += std::conj(Q[(size_t)r * m + i]) * SQ[(size_t)c * m + i];
        Sn[(size_t)c * k + r] = s;
      }
    FDFD_TRY(ops.rotate_basis(m, k, Q));
    std::fill(S.begin(), S.end(), cd(0, 0));
    for (int c = 0; c < k; ++c) {
      for (int r = 0; r < k; ++r) S[(size_t)c * ld + r] = Sn[(size_t)c * k + r];
      S[(size_t)c * ld + k] = bn[c];
    }
    m = k;
    ++out.restarts;
  }
  out.m = m;
  out.nu.resize(nev); out.resid.resize(nev); out.Y.assign((size_t)nev * m, cd(0, 0));
  for (int e = 0; e < nev; ++e) {
    out.nu[e] = evals[idx[e]];
    out.resid[e] = res[idx[e]];
    for (int i = 0; i < m; ++i) out.Y[(size_t)e * m + i] = evecs[(size_t)idx[e] * m + i];
  }
  return FDFD_OK;
}

// host-only test hook (no GPU needed): the whole Krylov-Schur loop -- expansion, Ritz test, thick restart, invariant-subspace
// handling -- on a dense n x n matrix OP (column major), vectors held in host memory.  out_nu[nev], out_vecs[nev][n]
extern "C" int fdfd_debug_krylov_schur(int n, const fdfd_c128* OP, int nev, int ncv, int which, double tol, int max_steps,
                                       fdfd_c128* out_nu, fdfd_c128* out_vecs, int* steps, int* restarts) {
  if (n < 2 || !OP || nev < 1 || !out_nu) return FDFD_ERR_ARG;
  const cd* A = reinterpret_cast<const cd*>(OP);
  std::vector<std::vector<cd>> V;
  std::vector<cd> w(n);
  ArnoldiOps ops;
  ops.op_apply = [&](int j) { for (int i = 0; i < n; ++i) { cd s = 0; for (int c = 0; c < n; ++c) s += A[(size_t)c * n + i] * V[j][c]; w[i] = s; } return FDFD_OK; };
  ops.dot_v_w = [&](int i, cd* h) { cd s = 0; for (int c = 0; c < n; ++c) s += std::conj(V[i][c]) * w[c]; *h = s; return FDFD_OK; };
  ops.axpy_w = [&](int i, cd h) { for (int c = 0; c < n; ++c) w[c] -= h * V[i][c]; return FDFD_OK; };
  ops.norm_w = [&](double* nr) { double s = 0; for (int c = 0; c < n; ++c) s += std::norm(w[c]); *nr = std::sqrt(s); return FDFD_OK; };
  ops.set_v = [&](int j, double s) { if ((int)V.size() <= j) V.resize(j + 1); V[j].resize(n); for (int c = 0; c < n; ++c) V[j][c] = s * w[c]; return FDFD_OK; };
  ops.random_w = [&](int seed) {
    uint64_t z = 0x9E3779B97F4A7C15ull * (uint64_t)(seed + 1);
    for (int c = 0; c < n; ++c) {
      z ^= z << 13; z ^= z >> 7; z ^= z << 17;
      w[c] = cd((double)(z & 0xFFFFFFFF) / 4294967296.0 - 0.5, (double)(z >> 32) / 4294967296.0 - 0.5);
    }
    return FDFD_OK;
  };
  ops.rotate_basis = [&](int m, int k, const std::vector<cd>& Q) {
    std::vector<std::vector<cd>> T(k, std::vector<cd>(n, cd(0, 0)));
    for (int i = 0; i < k; ++i) for (int j = 0; j < m; ++j) for (int c = 0; c < n; ++c) T[i][c] += Q[(size_t)i * m + j] * V[j][c];
    std::vector<cd> last = V[m];
    for (int i = 0; i < k; ++i) V[i] = T[i];
    V[k] = last;
    return FDFD_OK;
  };
  ArnoldiResult R;
  const int st = krylov_schur(nullptr, ops, nev, ncv, which, tol, max_steps, false, R);
  if (steps) *steps = R.steps;
  if (restarts) *restarts = R.restarts;
  if (st != FDFD_OK) return st;
  for (int e = 0; e < nev; ++e) {
    out_nu[e].re = R.nu[e].real(); out_nu[e].im = R.nu[e].imag();
    if (out_vecs) {
      for (int c = 0; c < n; ++c) {
        cd s = 0;
        for (int i = 0; i < R.m; ++i) s += R.Y[(size_t)e * R.m + i] * V[i][c];
        out_vecs[(size_t)e * n + c].re = s.real(); out_vecs[(size_t)e * n + c].im = s.imag();
      }
    }
  }
  return FDFD_OK;
}

extern "C" int fdfd_debug_general_eig(int n, const fdfd_c128* A, fdfd_c128* evals, fdfd_c128* evecs) {
  if (n < 1 || !A || !evals || !evecs) return FDFD_ERR_ARG;
  std::vector<cd> a((size_t)n * n), ev, vec;
  std::memcpy(a.data(), A, sizeof(cd) * n * n);
  if (!general_eig(n, a, ev, vec)) return FDFD_ERR_NOCONV;
  std::memcpy(evals, ev.data(), sizeof(cd) * n);
  std::memcpy(evecs, vec.data(), sizeof(cd) * n * n);
  return FDFD_OK;
}
