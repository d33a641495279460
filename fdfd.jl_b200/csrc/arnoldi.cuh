// arnoldi.cuh -- Krylov-Schur (thick-restart Arnoldi) driver shared by fdfd_eigenfrequency (eigen.cu) and
// fdfd_eigenfrequency_slab (slab_multi.cu).  Stands where Arpack's implicitly restarted Arnoldi stands in the reference
// (`eigs(A, nev=..., sigma=...)`, src/solver/eigen.jl:86,104): like Arpack, the basis never holds more than ncv + 1 vectors.
// Everything here is host code on the small projected matrix; the N-vectors are touched only through the callbacks, so the same
// loop drives a single-GPU basis, a row-slab sharded basis (dots allreduced) and the host-only test hook.
#pragma once
#include <complex>
#include <functional>
#include <vector>

struct fdfd_ctx;

struct ArnoldiOps {
  using cd = std::complex<double>;
  std::function<int(int j)> op_apply;                 // w = OP V[j]
  std::function<int(int i, cd* h)> dot_v_w;           // *h = <V[i], w>
  std::function<int(int i, cd h)> axpy_w;             // w -= h V[i]
  std::function<int(double* nrm)> norm_w;             // *nrm = ||w||
  std::function<int(int j, double s)> set_v;          // V[j] = s w   (V[j] is created if it does not exist yet)
  std::function<int(int seed)> random_w;              // w = pseudo-random vector (start vector, invariant-subspace restart)
  // V[i] <- sum_{j<m} Q[j + m i] V[j] for i < k, then V[k] <- V[m]   (Q is m x k, column major, orthonormal columns)
  std::function<int(int m, int k, const std::vector<cd>& Q)> rotate_basis;
};

struct ArnoldiResult {
  int m = 0;                                    // basis size the Ritz vectors refer to
  int steps = 0, restarts = 0;                  // operator applications, thick restarts
  std::vector<std::complex<double>> nu;         // nev Ritz values of OP, in `which` order
  std::vector<std::complex<double>> Y;          // nev coefficient vectors (m each): Ritz vector e = sum_i Y[e m + i] V[i]
  std::vector<double> resid;                    // their residual estimates |b^T y|
};

// which: FDFD_WHICH_* on the spectrum of OP (the shift-inverted spectrum, like ARPACK).  tol: |b^T y| <= tol |nu|.
int krylov_schur(fdfd_ctx* ctx, const ArnoldiOps& ops, int nev, int ncv, int which, double tol, int max_steps, bool verbose,
                 ArnoldiResult& out);

// eigen-decomposition of a small general complex matrix (column major n x n): Householder reduction to Hessenberg form,
// shifted QR (hess_eig), back-transformation; unit-norm eigenvectors in the columns of evecs
bool general_eig(int n, std::vector<std::complex<double>> A, std::vector<std::complex<double>>& evals,
                 std::vector<std::complex<double>>& evecs);
bool hess_eig(int n, std::vector<std::complex<double>> H, std::vector<std::complex<double>>& evals,
              std::vector<std::complex<double>>& evecs);
double which_key(int which, std::complex<double> nu);
