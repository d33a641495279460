// krylov.cuh -- GPU-resident right-preconditioned BiCGSTAB (K7), operator-agnostic: the driven TM/TE solve,
// the block-coupled sideband system of the modulated solve and the shift-invert inner solves of
// eigenfrequency all run through krylov_bicgstab with different apply / precondition callbacks.
#pragma once
#include "device_ops.cuh"
#include "mg.cuh"
#include <functional>
#include <map>

// device-resident solver scalars: no host round trip inside an iteration
struct KScal {
  c128 rho, rho_old, alpha, omega, beta;
  double bnorm2, rr, tol2;
  int done, breakdown, iter, pad;
};

struct IterGraph { cudaGraphExec_t exec = nullptr; std::vector<void*> post; int64_t nlaunch = 0; };

// vectors + scalars of one Krylov solve, resident in HBM
struct KrylovWork {
  int64_t n = 0;
  DevBuf<c128> b, x, r, rhat, p, v, s, t;
  DevBuf<c128> ph, sh;      // fp64 preconditioned vectors (Jacobi only)
  DevBuf<c128> partials;    // per-CTA partial sums, [block][<=2] complex
  DevBuf<KScal> scal;
  DevBuf<double> lsum;      // slab mode: locally reduced sums handed to the allreduce
  DevBuf<double> hist;      // ||r||^2 per iteration
  KScal* h_scal = nullptr;  // pinned mirror
  int nvec_blocks = 0;
  std::map<std::vector<void*>, IterGraph> graphs;  // one captured iteration per buffer-rotation state
  int alloc(fdfd_ctx* ctx, int64_t n, int nparts, int maxit, bool jacobi_bufs);
  ~KrylovWork();
};

// what the solver needs to know about the system
struct KrylovOps {
  bool prec_f32 = true;        // preconditioned vectors are c64 (fp32 multigrid) else c128
  void* prec_rhs = nullptr;    // update kernels write the (scaled) copy of p / s here (multigrid rhs) or null
  double fscale = 1.0;         // scale applied to that copy
  int nab = 0;                 // partial blocks produced by apply
  // y = A x with fused dots; x is c64 when x_f32
  std::function<int(const void* x, bool x_f32, c128* y, const DotSpec& ds)> apply;
  // result of M^-1 applied to the vector previously written to prec_rhs (or to p / s when prec_rhs == null)
  std::function<int(bool hold, const void** out)> precond;
  std::function<int(double* dev4)> allreduce;                    // slab mode: sum 4 doubles over the ranks (in place, on the stream)
  std::function<void(std::vector<void*>&)> get_state;            // buffer-rotation state (may be empty)
  std::function<void(const std::vector<void*>&)> set_state;
};

int krylov_bicgstab(fdfd_ctx* ctx, KrylovWork& W, const KrylovOps& ops, const fdfd_solve_opts_t& o, fdfd_info_t* info);
int vec_blocks_for(fdfd_ctx* ctx, int64_t n);
int apply_num_blocks(int64_t nx, int64_t ny);

struct MLKrylov;
void mlkrylov_free(MLKrylov* m);

struct fdfd_problem {
  fdfd_ctx* ctx = nullptr;
  FineOp op;
  fdfd_solve_opts_t opts{};
  Multigrid<float>* mgf = nullptr;
  Multigrid<double>* mgd = nullptr;
  KrylovWork w;
  double setup_ms = 0;
  bool have_rhs = false;
  MLKrylov* ml = nullptr;   // state of the multilevel Krylov solver (FDFD_SOLVER_MLKRYLOV), built at the first solve
  KrylovOps make_ops();
  ~fdfd_problem();
};

MGParams mg_params_from(const fdfd_solve_opts_t& o);
int krylov_cocg(fdfd_problem* P, fdfd_info_t* info);
int krylov_multilevel(fdfd_problem* P, fdfd_info_t* info);
int jacobi_apply(fdfd_ctx* ctx, const FineOp& op, const c128* in, c128* out, const int* done, int blocks);
