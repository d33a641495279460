// krylov.cuh -- GPU-resident right-preconditioned BiCGSTAB (K7) and the resident problem handle.
#pragma once
#include "device_ops.cuh"
#include "mg.cuh"

// device-resident solver scalars: no host round trip inside an iteration
struct KScal {
  c128 rho, rho_old, alpha, omega, beta;
  double bnorm2, rr, tol2;
  int done, breakdown, iter, pad;
};

struct fdfd_problem {
  fdfd_ctx* ctx = nullptr;
  FineOp op;
  fdfd_solve_opts_t opts{};
  Multigrid<float>* mgf = nullptr;
  Multigrid<double>* mgd = nullptr;
  DevBuf<c128> b, x, r, rhat, p, v, s, t;
  DevBuf<c128> ph, sh;      // fp64 preconditioned vectors (Jacobi / none)
  DevBuf<c128> partials;    // [blocks][<=2] complex partial sums
  DevBuf<KScal> scal;
  DevBuf<double> hist;      // ||r||^2 per iteration
  KScal* h_scal = nullptr;  // pinned mirror
  int nvec_blocks = 0;
  double setup_ms = 0;
  bool have_rhs = false, have_x = false;
  ~fdfd_problem();
};

int problem_solve_bicgstab(fdfd_problem* p, fdfd_info_t* info);
