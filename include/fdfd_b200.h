/* fdfd_b200.h -- C ABI of the B200-native FDFD.jl assembly + linear-solve hot path.
 *
 * Drop-in boundary.  The reference (fancompute/FDFD.jl, pure Julia) has
 * one plug-in seam, `dolinearsolve(A, b, sym)` (src/solver/solver.jl:4), but assembly is
 * part of the path and the GPU solver is matrix-free, so the entry points below replace the
 * L3 functions one level higher; each cites the reference code it stands in for.  A Julia
 * maintainer binds them with `ccall`; tests bind them with ctypes.
 *
 * Conventions
 *   - complex128 == two adjacent doubles (Julia ComplexF64 / C double _Complex / double2).
 *   - all grid arrays are column-major (Nx,Ny) with x fastest: n = ix + Nx*iy  (Julia `a[:]`).
 *   - field outputs are (Nx,Ny,3) column-major, components [Ez,Hx,Hy] (TM) / [Hz,Ex,Ey] (TE)
 *     exactly like FieldTM/FieldTE.data (src/data.jl:50-78).
 *   - every buffer is caller-owned; pointers may be host (pageable or pinned) or device
 *     pointers (detected with cudaPointerGetAttributes).  Nothing is retained after return
 *     except inside an explicit fdfd_problem handle, which copies what it needs.
 *   - every function returns 0 on success, non-zero on error (message: fdfd_last_error).
 *     No C++ exception crosses the boundary.  There is NO CPU fallback: without a CUDA
 *     device every compute entry point fails with FDFD_ERR_CUDA.
 */
#ifndef FDFD_B200_H
#define FDFD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDFD_B200_ABI_VERSION 1

/* status codes */
enum {
  FDFD_OK = 0,
  FDFD_ERR_ARG = 1,      /* bad argument */
  FDFD_ERR_CUDA = 2,     /* CUDA runtime error / no device */
  FDFD_ERR_NOCONV = 3,   /* Krylov solver hit maxit (results still written, see info) */
  FDFD_ERR_BREAKDOWN = 4,/* Krylov breakdown that restarts could not cure */
  FDFD_ERR_ALLOC = 5
};

/* src/types.jl:20 */
enum { FDFD_TM = 1, FDFD_TE = 2 };
/* operator ordering: f.b = Dxf*Dxb (src/solver/driven.jl:35, eigen.jl:84,102),
 *                    b.f = Dxb*Dxf (src/solver/modulation.jl:82) */
enum { FDFD_ORDER_FB = 0, FDFD_ORDER_BF = 1 };
/* derivative selector for fdfd_assemble_derivative (src/grid.jl:128-154) */
enum { FDFD_DXF = 0, FDFD_DXB = 1, FDFD_DYF = 2, FDFD_DYB = 3 };
/* sparse formats: CSR (row pointers) or CSC (Julia SparseMatrixCSC: colptr,rowval,nzval) */
enum { FDFD_CSR = 0, FDFD_CSC = 1 };
/* Krylov solvers / preconditioners */
/* BICGSTAB: any preconditioner.  COCG: on the symmetrised system diag(sxf*syf) A, Jacobi or no preconditioner. */
/* MLKRYLOV (csrc/mlkrylov.cu): multilevel Krylov -- flexible GMRES on every level, preconditioned by the multigrid cycle plus a
 * coarse-grid Helmholtz correction solved by the same method one level down.  TM and TE; FDFD_PRECOND_MG and FDFD_MG_F32 only.
 * AUTO (default): MLKRYLOV for the driven single-GPU solve of a grid of >= 2^22 points with the fp32 multigrid (measured on a
 * B200, 4096^2 bench map: 105 outer iterations / 3.3-4.1 s against 1230 BiCGSTAB iterations / 5.0 s; below 2048^2 both are
 * launch-latency bound and BiCGSTAB is as fast), BICGSTAB everywhere else (slab-sharded, modulated, Jacobi / no preconditioner). */
enum { FDFD_SOLVER_BICGSTAB = 0, FDFD_SOLVER_COCG = 1, FDFD_SOLVER_MLKRYLOV = 2, FDFD_SOLVER_AUTO = 3 };
enum { FDFD_PRECOND_NONE = 0, FDFD_PRECOND_JACOBI = 1, FDFD_PRECOND_MG = 2 };
enum { FDFD_MG_F32 = 0, FDFD_MG_F64 = 1 };
enum { FDFD_CYCLE_V = 0, FDFD_CYCLE_F = 1, FDFD_CYCLE_W = 2 };
/* eigenfrequency `which` (Arpack semantics on the shift-inverted spectrum, eigen.jl:86) */
enum { FDFD_WHICH_LM = 0, FDFD_WHICH_LR = 1, FDFD_WHICH_SR = 2, FDFD_WHICH_LI = 3, FDFD_WHICH_SI = 4 };

typedef struct { double re, im; } fdfd_c128;

/* POD mirror of Grid{2} (src/grid.jl:7-13); N is computed by the caller exactly as the
 * reference constructor does (round(L/dh), src/grid.jl:28).  dx = (x1-x0)/Nx (grid.jl:68). */
typedef struct {
  int64_t Nx, Ny;
  int64_t Npml_x, Npml_y;
  double x0, x1, y0, y1;
  double L0;
} fdfd_grid_t;

typedef struct {
  int32_t solver;        /* FDFD_SOLVER_*  (default AUTO) */
  int32_t precond;       /* FDFD_PRECOND_* (default MG) */
  double  tol;           /* relative residual ||b-Ax||/||b|| of the un-preconditioned system (default 1e-10) */
  int32_t maxit;         /* Krylov iterations (default 20000) */
  int32_t mg_precision;  /* FDFD_MG_F32 | FDFD_MG_F64 (default F32) */
  int32_t mg_cycle;      /* FDFD_CYCLE_* of the BiCGSTAB preconditioner (default W, truncated at mg_wdepth).  The multilevel Krylov solver
                            always runs F cycles for its M_l^-1 (a W cycle truncated at an absolute depth degenerates to V on its inner
                            levels; measured 105 against 210 outer iterations at 4096^2) */
  int32_t mg_wdepth;     /* levels [0,wdepth) recurse twice in a W cycle (default 3; measured at 4096^2: 1982 / 1230 / 1263 BiCGSTAB
                            iterations and 6.6 / 5.0 / 6.5 s for depth 2 / 3 / 4) */
  int32_t mg_nu1, mg_nu2;/* pre/post smoothing sweeps (default 1,1) */
  int32_t mg_coarse_sweeps; /* sweeps on the coarsest level (default 2) */
  double  mg_beta;       /* complex shift: M = L + (1 - i*beta) w^2 eps (default 0.5) */
  double  mg_wjac;       /* point-Jacobi damping (default 0.7; measured on the 4096^2 sweep: 0.8 needs 15 % more iterations, 0.9 twice as many) */
  double  mg_wline;      /* PML line-relaxation damping (default 0.6) */
  int32_t check_every;   /* host polls convergence every k iterations (default 8) */
  int32_t verbose;
  double  mg_shift_growth; /* level/space dependent shift: beta_eff = max(beta, growth * Re(k^2 h_l^2)) (default 0 = off) */
  int32_t mg_max_levels; /* cap on hierarchy depth (default 32) */
  int32_t use_graph;     /* replay the iteration as a CUDA graph (default 1) */
  int32_t concurrency;   /* fdfd_solve_driven: frequencies solved concurrently on separate streams (default 4) */
  int32_t ml_spec;       /* FDFD_SOLVER_MLKRYLOV: k1 | k2<<8 | k3<<16 | restart<<24 = FGMRES steps per solve on levels 1,2,3 (0 ends the
                            list) and the level-0 restart length; 0 = defaults (6, 6; restart 96, or what a 1/concurrency share of the free device
                            memory holds inside fdfd_solve_driven; the level-0 basis grows on demand, 2 vectors per iteration, and a full
                            device forces an earlier restart instead of an error) */
} fdfd_solve_opts_t;

typedef struct {
  int32_t iters;         /* Krylov iterations used */
  int32_t flag;          /* FDFD_OK | FDFD_ERR_NOCONV | FDFD_ERR_BREAKDOWN */
  double  relres;        /* final TRUE relative residual, recomputed with the fp64 operator */
  double  setup_ms;      /* coefficient + MG hierarchy build (device) */
  double  solve_ms;      /* Krylov loop, CUDA events on the ctx stream */
  double  total_ms;      /* wall time of the call incl. host<->device copies */
  int64_t launches;      /* kernels launched by this call */
  int32_t restarts;
  int32_t mg_levels;
} fdfd_info_t;

typedef struct fdfd_ctx fdfd_ctx;
typedef struct fdfd_problem fdfd_problem;

/* ---- context -------------------------------------------------------------------------- */
int fdfd_abi_version(void);
/* device: CUDA ordinal.  stream: an existing cudaStream_t to launch on (e.g. torch's current
 * stream) or NULL to let the library create its own non-blocking stream. */
int fdfd_ctx_create(int device, void* stream, fdfd_ctx** out);
void fdfd_ctx_destroy(fdfd_ctx* ctx);
const char* fdfd_last_error(fdfd_ctx* ctx);   /* ctx may be NULL: last global error */
int64_t fdfd_launch_count(fdfd_ctx* ctx);     /* kernels launched since ctx creation */
void fdfd_default_opts(fdfd_solve_opts_t* opts);

/* ---- K1: PML s-factors.  Replaces create_sfactor (src/pml.jl:1-31): writes the four 1-D
 * arrays s (NOT inverted), lengths Nx,Nx,Ny,Ny, order (x fwd, x bwd, y fwd, y bwd). */
int fdfd_sfactors(fdfd_ctx* ctx, const fdfd_grid_t* g, double omega,
                  fdfd_c128* sxf, fdfd_c128* sxb, fdfd_c128* syf, fdfd_c128* syb);

/* ---- K2: Yee derivative operators assembled on the GPU straight into CSR/CSC.
 * Replaces δ(w,s,g) (src/grid.jl:128-154) and, with stretched!=0, the row scaling by the
 * inverse s-factors `Sxb*δ(...)` (src/pml.jl:33-63, src/solver/driven.jl:28-31).
 * ptr has N+1 entries, ind/val have 2N entries; index_base is 0 or 1 (Julia). */
int fdfd_assemble_derivative(fdfd_ctx* ctx, const fdfd_grid_t* g, double omega, int which,
                             int stretched, int format, int index_base,
                             int64_t* ptr, int64_t* ind, fdfd_c128* val);

/* ---- K3: system matrix assembled on the GPU into CSR/CSC (5 nnz per row).
 * pol=TM, FB : A = Dxf*mu0^-1*Dxb + Dyf*mu0^-1*Dyb + w^2*diag(eps0*eps_r)   (driven.jl:35)
 * pol=TM, BF : A = Dxb/mu0*Dxf + Dyb/mu0*Dyf + w^2*diag(eps0*eps_r)         (modulation.jl:82,85)
 * pol=TE, FB : A = Dxf*diag(1/avgx)*Dxb + Dyf*diag(1/avgy)*Dyb + w^2*mu0*I  (driven.jl:45)
 * ptr: N+1, ind/val: 5N. */
int fdfd_assemble_system(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, int ordering, double omega,
                         const fdfd_c128* eps_r, int format, int index_base,
                         int64_t* ptr, int64_t* ind, fdfd_c128* val);

/* ---- K4/K5: matrix-free operator apply y = A x (same A as fdfd_assemble_system), host or
 * device buffers.  Kernel-level parity hook. */
int fdfd_apply_operator(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, int ordering, double omega,
                        const fdfd_c128* eps_r, const fdfd_c128* x, fdfd_c128* y);

/* ---- driven solve.  Replaces solve(d::Device, pol) (src/solver/driven.jl:4-59) for n_omega
 * frequencies sharing eps_r; src is (Nx,Ny) per frequency when src_per_omega!=0 else shared.
 * b = i*w*src (driven.jl:36).  fields: n_omega x (Nx,Ny,3).  info: n_omega entries. */
int fdfd_solve_driven(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, int n_omega, const double* omega,
                      const fdfd_c128* eps_r, const fdfd_c128* src, int src_per_omega,
                      const fdfd_solve_opts_t* opts, fdfd_c128* fields, fdfd_info_t* info);

/* ---- modulated multi-frequency solve.  Replaces solve(d::ModulatedDevice)
 * (src/solver/modulation.jl:35-119) for one w: nf = 2*nsidebands+1 coupled sidebands,
 * fields: nf x (Nx,Ny,3) (sideband -ns first), H from forward differences (:112-113). */
int fdfd_solve_modulated(fdfd_ctx* ctx, const fdfd_grid_t* g, double omega, double Omega, int nsidebands,
                         int sharedpml, const fdfd_c128* eps_r, const fdfd_c128* deps_r,
                         const fdfd_c128* src, const fdfd_solve_opts_t* opts,
                         fdfd_c128* fields, fdfd_info_t* info);

/* ---- eigenfrequency.  Replaces eigenfrequency(d, pol, nev; which) (src/solver/eigen.jl:69-115):
 * shift-invert Arnoldi around sigma = -w0^2 mu0 eps0 (TM) / -w0^2 mu0 (TE) with the PML frozen
 * at w0; the inner solves reuse the driven operator.  omega_out: nev complex; fields: nev x (Nx,Ny,3)
 * (may be NULL).  ncv<=0 picks max(20, 2*nev+1) like Arpack.jl, and like Arpack the basis never holds more than ncv + 1
 * vectors: Krylov-Schur thick restarts (csrc/arnoldi.cu) keep the nev + (ncv - nev)/2 best Ritz vectors.  A Ritz pair counts as
 * converged at |b^T y| <= 10 * min(opts->tol, 1e-11) * |nu| (Arpack's machine-epsilon default presumes an exact factorisation
 * behind the operator; here it is an iterative solve to min(opts->tol, 1e-11)); at most max(300, 30 nev) + ncv operator
 * applications, then FDFD_ERR_NOCONV.  info->restarts reports the operator applications, info->iters the inner Krylov iterations. */
int fdfd_eigenfrequency(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, double omega0, int nev, int which,
                        int ncv, const fdfd_c128* eps_r, const fdfd_solve_opts_t* opts,
                        fdfd_c128* omega_out, fdfd_c128* fields, fdfd_info_t* info);

/* ---- resident-problem handle (what the one-shot calls are built from; lets a sweep keep
 * eps_r, coefficients and the multigrid hierarchy in HBM between solves). */
int fdfd_problem_create(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, int ordering, double omega,
                        const fdfd_c128* eps_r, const fdfd_solve_opts_t* opts, fdfd_problem** out);
void fdfd_problem_destroy(fdfd_problem* p);
int fdfd_problem_set_rhs(fdfd_problem* p, const fdfd_c128* b);            /* b as is */
int fdfd_problem_set_source(fdfd_problem* p, const fdfd_c128* src);       /* b = i*w*src */
int fdfd_problem_solve(fdfd_problem* p, fdfd_info_t* info);               /* device-resident */
int fdfd_problem_get_solution(fdfd_problem* p, fdfd_c128* x);             /* (Nx,Ny) */
int fdfd_problem_get_fields(fdfd_problem* p, int forward_h, fdfd_c128* fields); /* (Nx,Ny,3) */
/* timed loop of nrep matrix-free applies on resident data; ms_per_apply from CUDA events */
int fdfd_problem_bench_apply(fdfd_problem* p, int nrep, double* ms_per_apply);
/* timed loop of one multigrid kernel on the level-0 arrays of the resident hierarchy (fp32 multigrid only).  kind: 0 = smoothing
 * sweep fused with the coarse-grid correction (k_smooth3 + the PML line kernel), 1 = residual + restriction
 * (k_restrict_tile), 2 = zero-guess sweep (k_smooth2 + the PML line kernel), 4 = one whole cycle (M^-1 applied once). */
int fdfd_problem_bench_mg(fdfd_problem* p, int kind, int nrep, double* ms_per_launch);
/* relative residual history of the last solve: out[k] = ||r_k|| / ||b|| (recurrence residual), k = 0..n-1;
 * returns the number of entries written through *written */
int fdfd_problem_get_history(fdfd_problem* p, double* out, int n, int* written);
/* multigrid cycles started on levels 0..3 by the last FDFD_SOLVER_MLKRYLOV solve (diagnostics) */
int fdfd_problem_ml_cycles(fdfd_problem* p, int64_t* out4);
/* one application of the preconditioner M^-1 to a resident vector (parity/debug hook) */
int fdfd_problem_precond(fdfd_problem* p, const fdfd_c128* in, fdfd_c128* out);

/* ---- callers either side of the path, kept on the device for large sweeps -------------------------
 * flux_surface_integral(field, Point(center_x, center_y), width, x̂) of the TM solution resident in `p`
 * (src/flux.jl:37-47); forward_h selects the H recovery of modulation.jl:112-113 / eigen.jl:90-91 instead of
 * driven.jl:40-41.  One double comes back instead of the (Nx,Ny,3) field. */
int fdfd_problem_flux_x(fdfd_problem* p, double center_x, double center_y, double width, int forward_h, double* flux);
/* setup_ϵᵣ!(d, shapes) (src/device.jl:47-61) for boxes/cylinders: shapes7 = nshapes x {kind(0 box,1 cylinder), cx, cy,
 * a, b, eps_re, eps_im} (box: a,b full widths; cylinder: a radius); the first shape containing a pixel centre wins,
 * other pixels keep the value already in eps_r (host or device buffer, (Nx,Ny)). */
int fdfd_rasterize(fdfd_ctx* ctx, const fdfd_grid_t* g, int nshapes, const double* shapes7, fdfd_c128* eps_r);

/* ---- one large grid split into row slabs (BASELINE config 5) -------------------------------------------
 * The reference has no parallel path at all (its solve is one `lu(A)\b`, src/solver/solver.jl:35); this is the sharded
 * form of solve(d::Device, TM) (src/solver/driven.jl:4-59) for grids that do not fit / are too slow on one GPU.
 * The global grid is cut into `nranks` contiguous y-slabs of Ny/nranks rows (x stays the fast, contiguous index); slab
 * r owns global rows [r*Ny/nranks, (r+1)*Ny/nranks).  Each slab runs the same BiCGSTAB + multigrid on its rows; the
 * only exchanges are (1) ring halo rows between neighbouring slabs (the operators are periodic, grid.jl:147,150, so the
 * ring closes) after every stencil-type kernel and (2) one sum of <= 4 doubles over the ranks per Krylov dot product.
 *
 * Communicators.  FDFD_COMM_NCCL: one process per GPU; rank 0 calls fdfd_comm_unique_id, the host program broadcasts
 * the FDFD_COMM_ID_BYTES bytes (torch.distributed / MPI / a file) and every rank calls fdfd_comm_create_nccl.
 * FDFD_COMM_THREADS: all slabs inside one process, one host thread per slab (any mix of GPUs, including all on one
 * GPU): fdfd_comm_group_create once, fdfd_comm_create_threads per rank.  Calls on a communicator are collective:
 * every rank must make the same calls in the same order. */
enum { FDFD_COMM_THREADS = 0, FDFD_COMM_NCCL = 1 };
#define FDFD_COMM_ID_BYTES 128
typedef struct fdfd_comm fdfd_comm;
typedef struct fdfd_comm_group fdfd_comm_group;
int fdfd_comm_unique_id(void* id_bytes);
int fdfd_comm_create_nccl(fdfd_ctx* ctx, int nranks, int rank, const void* id_bytes, fdfd_comm** out);
int fdfd_comm_group_create(int nranks, fdfd_comm_group** out);
void fdfd_comm_group_destroy(fdfd_comm_group* grp);
int fdfd_comm_create_threads(fdfd_comm_group* grp, int rank, fdfd_comm** out);
void fdfd_comm_destroy(fdfd_comm* comm);
/* rows owned by `rank` of `nranks`: *y0 = first global row, *nrows = Ny/nranks.  Host-only helper (no GPU needed). */
int fdfd_slab_rows(const fdfd_grid_t* g, int nranks, int rank, int64_t* y0, int64_t* nrows);
/* the solve.  g is the GLOBAL grid; eps_r_rows, src_rows: this rank's owned rows, (Nx, nrows) x fastest (a contiguous
 * row range of the global column-major arrays); fields_rows: (Nx, nrows, 3).  info is identical on every rank except
 * for the timings.  Ny must be divisible by nranks and the slab height by 2^(multigrid levels - 1); the hierarchy is
 * truncated to the deepest depth that allows it. */
int fdfd_solve_driven_slab(fdfd_ctx* ctx, fdfd_comm* comm, const fdfd_grid_t* g, double omega,
                           const fdfd_c128* eps_r_rows, const fdfd_c128* src_rows, const fdfd_solve_opts_t* opts,
                           fdfd_c128* fields_rows, fdfd_info_t* info);
/* ---- slab-sharded modulated and eigenfrequency solves (csrc/slab_multi.cu).
 * Same slab layout, communicators and collective-call rules as fdfd_solve_driven_slab.
 *
 * solve(d::ModulatedDevice) (src/solver/modulation.jl:35-119) on row slabs: the sideband coupling is pointwise
 * (modulation.jl:95-98), so all nf = 2*nsidebands+1 sidebands of a row live on the rank that owns the row and the
 * coupling needs no exchange.  eps_r_rows, deps_r_rows, src_rows: (Nx, nrows); fields_rows: nf x (Nx, nrows, 3),
 * sideband -ns first, H from forward differences (modulation.jl:112-113). */
int fdfd_solve_modulated_slab(fdfd_ctx* ctx, fdfd_comm* comm, const fdfd_grid_t* g, double omega, double Omega,
                              int nsidebands, int sharedpml, const fdfd_c128* eps_r_rows, const fdfd_c128* deps_r_rows,
                              const fdfd_c128* src_rows, const fdfd_solve_opts_t* opts, fdfd_c128* fields_rows,
                              fdfd_info_t* info);
/* eigenfrequency(d, TM, nev; which) (src/solver/eigen.jl:69-96) on row slabs: the Arnoldi basis is sharded like every
 * other vector, the inner shift-invert solves are slab solves, the Hessenberg matrix is replicated on the hosts.
 * pol must be FDFD_TM.  omega_out: nev complex (identical on every rank); fields_rows: nev x (Nx, nrows, 3) or NULL. */
int fdfd_eigenfrequency_slab(fdfd_ctx* ctx, fdfd_comm* comm, const fdfd_grid_t* g, int pol, double omega0, int nev,
                             int which, int ncv, const fdfd_c128* eps_r_rows, const fdfd_solve_opts_t* opts,
                             fdfd_c128* omega_out, fdfd_c128* fields_rows, fdfd_info_t* info);
/* counters of the communicator since creation: exchanges, allreduces, bytes sent by this rank */
int fdfd_comm_stats(fdfd_comm* comm, int64_t* n_exchange, int64_t* n_allreduce, int64_t* bytes_sent);

/* ---- dolinearsolve-level entry.  Replaces dolinearsolve(A::SparseMatrixCSC, b, matrixsym) -> x (src/solver/solver.jl:4-41)
 * for callers that assemble their own matrix and only use the solver seam: the chi-3 outer loops (src/solver/nonlinear.jl:69,97,120)
 * and the 2-D eigenmode (src/solver/eigen.jl:32-66).  colptr (n+1), rowval, nzval are HOST arrays as Julia stores them
 * (A.colptr, A.rowval, A.nzval with index_base 1; 0 for C callers); duplicates are summed; b, x (n) host or device.
 * Solver: BiCGSTAB + Jacobi over a SELL-32 image of A (a matrix carries no grid, so no multigrid): a compatibility path, the
 * driven / modulated / eigenfrequency entry points above are the fast ones.  opts: tol, maxit, check_every, use_graph, verbose. */
int fdfd_dolinearsolve_csc(fdfd_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval, const fdfd_c128* nzval,
                           int index_base, const fdfd_c128* b, const fdfd_solve_opts_t* opts, fdfd_c128* x, fdfd_info_t* info);
/* The same seam with the grid the matrix was assembled on (every caller of dolinearsolve has one, nonlinear.jl:58-66).  If A IS the
 * TM operator of g at omega for some permittivity -- the first solve (nonlinear.jl:66-69) and every Born step
 * A + Diagonal(coeff |ez|^2) (nonlinear.jl:97) are -- eps_eff = (A 1)/(w^2 eps0 L0) is read off the row sums, A v == A_matrixfree v
 * is checked on the device (relative 1e-9, both derivative orderings) and the multigrid-preconditioned solver runs
 * (info->mg_levels > 0); anything else (the 2N x 2N Gauss-Newton Jacobian, nonlinear.jl:120; a TE matrix) takes the generic path of
 * fdfd_dolinearsolve_csc (info->mg_levels == 0).  info->relres is recomputed against the caller's matrix. */
int fdfd_dolinearsolve_csc_grid(fdfd_ctx* ctx, const fdfd_grid_t* g, double omega, int64_t n, const int64_t* colptr,
                                const int64_t* rowval, const fdfd_c128* nzval, int index_base, const fdfd_c128* b,
                                const fdfd_solve_opts_t* opts, fdfd_c128* x, fdfd_info_t* info);
/* measurement hook (GPU): average ms of `reps` launches of the SELL-32 SpMV with two fused dots (CUDA events on the ctx stream,
 * 3 warm-up launches) and the algorithmic bytes of one launch (36 B per stored entry + 32 B per row). */
int fdfd_debug_sell_bench(fdfd_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval, const fdfd_c128* nzval,
                          int index_base, int reps, double* ms_per_launch, double* alg_bytes);
/* host-only test hook (no GPU needed): y = A x through the same CSC -> SELL-32 transposition and per-row summation order as the
 * SpMV kernel; dinv (n, optional) = the Jacobi preconditioner's inverse diagonal; rowsum (n, optional) = A 1, from which the
 * grid-hinted path reads the permittivity; padded_entries (optional) = stored entries. */
int fdfd_debug_sell_spmv(int64_t n, const int64_t* colptr, const int64_t* rowval, const fdfd_c128* nzval, int index_base,
                         const fdfd_c128* x, fdfd_c128* y, fdfd_c128* dinv, fdfd_c128* rowsum, int64_t* padded_entries);

/* host-only test hook (no GPU needed): the small dense complex Hessenberg eigen-solver behind the Ritz pairs of
 * fdfd_eigenfrequency.  H, evecs: column-major n x n; evals: n. */
/* y_b = A x_b for nrhs right-hand sides that share ONE operator (sources sharing a frequency -- the batch axis of the reference's sweep
 * loop, driven.jl:11): the stencil reads the coefficients, above all w^2 eps, once per point for all nrhs vectors:
 * (32 nrhs + 16) / nrhs algorithmic bytes per point and right-hand side instead of 48.  Equal to nrhs single
 * fdfd_apply_operator calls to rounding (same arithmetic, different FMA contraction).  x, y: nrhs x (Nx,Ny) one after the other.  Kernel-level building block and benchmark: the Krylov
 * solvers still take one right-hand side per solve (concurrent streams). */
int fdfd_apply_operator_batched(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, int ordering, double omega,
                                const fdfd_c128* eps_r, int nrhs, const fdfd_c128* x, fdfd_c128* y);
int fdfd_problem_bench_apply_batched(fdfd_problem* p, int nrhs, int nrep, double* ms_per_launch);

int fdfd_debug_hess_eig(int n, const fdfd_c128* H, fdfd_c128* evals, fdfd_c128* evecs);
/* host-only test hooks (no GPU needed) of the Krylov-Schur driver behind fdfd_eigenfrequency[_slab] (csrc/arnoldi.cu; stands
 * where Arpack's implicitly restarted Arnoldi stands, eigen.jl:86,104): the eigen-solver of a general small matrix (the projected
 * matrix is no longer Hessenberg after a thick restart), and the whole loop -- expansion to ncv, Ritz test |b^T y| <= tol |nu|,
 * thick restart, invariant-subspace handling -- on a dense n x n operator OP (column major) with host vectors.
 * `which` = FDFD_WHICH_* on OP's spectrum; out_nu[nev]; out_vecs (may be NULL): nev x n Ritz vectors. */
int fdfd_debug_general_eig(int n, const fdfd_c128* A, fdfd_c128* evals, fdfd_c128* evecs);
int fdfd_debug_krylov_schur(int n, const fdfd_c128* OP, int nev, int ncv, int which, double tol, int max_steps,
                            fdfd_c128* out_nu, fdfd_c128* out_vecs, int* steps, int* restarts);

/* host-only test hooks of FDFD_SOLVER_MLKRYLOV (no GPU needed): the one-thread least-squares solve of the (k+1) x k
 * Hessenberg system min || beta e1 - H y || (H column-major, ld = k+1), and the grid transfers between a fine (nx,ny) grid
 * and its vertex-centred coarse grid ((nx+1)/2, (ny+1)/2): mode 0 = scale * Z^T (fine -> coarse), mode 1 = Z (coarse -> fine). */
int fdfd_debug_ml_lsq(int k, const fdfd_c128* H, double beta, fdfd_c128* y, double* resnorm);
int fdfd_debug_ml_transfer(int64_t nx, int64_t ny, int mode, double scale, const fdfd_c128* in, fdfd_c128* out);
/* the same two cores run by their device kernels (GPU needed; host or device buffers) -- device/host comparison hooks */
int fdfd_debug_ml_lsq_gpu(fdfd_ctx* ctx, int k, const fdfd_c128* H, double beta, fdfd_c128* y, double* resnorm);
int fdfd_debug_ml_transfer_gpu(fdfd_ctx* ctx, int64_t nx, int64_t ny, int mode, double scale, const fdfd_c128* in, fdfd_c128* out);

#ifdef __cplusplus
}
#endif
#endif /* FDFD_B200_H */
