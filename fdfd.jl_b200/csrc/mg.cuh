// mg.cuh -- matrix-free geometric multigrid on the complex-shifted operator
//   M = L_pml + (1 - i beta) w^2 eps        (TM; TE shifts the constant w^2 mu0 term)
// used as the right preconditioner of the Krylov solver (K8).  New work: the reference has no
// iterative solver (src/solver/solver.jl:29-35 is a sparse direct LU).
//
// Design (calibrated in tools/mg_prototype.py):
//  * vertex-centred coarsening  coarse I <-> fine 2I, any size (odd sizes leave one short coarse edge).
//    Transfers and coarse operators are Galerkin-consistent in the STRETCHED coordinate: the PML makes the
//    operator a variable-(complex)-coefficient one, so interpolation is linear in x~ = int s dx (weights from
//    the edge conductances), restriction is its transpose weighted by the stretched cell volumes, and the coarse
//    1-D PML coefficients are built recursively (coarse edge = sum of fine edges, coarse volume = P^T V).
//    Plain bilinear/full-weighting + point-sampled s-profiles diverge when the PML is thinner than a coarse cell
//    (measured: dh = 0.01, Npml = 10: cycle factor 2.7 -> 0.33 with the consistent transfers).
//  * smoother: damped point Jacobi in the interior; inside the PML strips the stretched operator is
//    strongly anisotropic with rotated phases and point smoothers amplify, so the strip columns get
//    y-line relaxation and the strip rows x-line relaxation.  Line systems are solved by parallel cyclic
//    reduction with multipliers precomputed at setup (only the right-hand side is reduced per sweep).
//  * arithmetic in T = float (default; halves HBM traffic, the outer Krylov is fp64) or double.
#pragma once
#include "device_ops.cuh"

// A PML line (a strip column along y or a strip row along x) is relaxed in overlapping segments: segment j solves the
// tridiagonal system of line points [j Lc - O, (j+1) Lc + O) (couplings cut at its ends) and updates its core
// [j Lc, (j+1) Lc).  Inside the strips the line operator is diagonally dominant (|rho| <= ~0.95 per cell in the deepest
// PML cells at level 0, far smaller elsewhere), so a cut O >= 128 cells away changes the core by < 1e-2 of a smoothing
// update, while one 4096-point line becomes 16 independent 512-point systems: 1024 CTAs instead of 64 and 9 PCR steps
// instead of 12 (measured: 67 us -> see DESIGN.md).  Lines of <= 640 points are a single exact segment.
struct LineSeg {
  int n = 0, SL = 0, Lc = 0, O = 0, nseg = 1, K = 0;   // line length, stored segment length, core, overlap, segments, PCR steps
  __host__ __device__ int lo(int j) const { const int a = j * Lc - O; return a < 0 ? 0 : a; }
  __host__ __device__ int hi(int j) const { const int b = (j + 1) * Lc + O; return b > n ? n : b; }
  __host__ __device__ int core_lo(int j) const { return j * Lc; }
  __host__ __device__ int core_hi(int j) const { const int b = (j + 1) * Lc; return (b > n || j == nseg - 1) ? n : b; }
};

// rows of the y-direction PML strip as two ranges [a0,a1) U [b0,b1)
struct YS { int a0 = 0, a1 = 0, b0 = 0, b1 = 0; int count() const { return (a1 - a0) + (b1 - b0); } };

template <typename T> struct MGLevel {
  int64_t nx = 0, ny = 0, stride = 1;
  YS ys;                      // y-strip rows of this level (single GPU: [0,npy) U [ny-npy,ny))
  int npx = 0, npy = 0;       // strip half-widths: columns [0,npx) U [nx-npx,nx), rows likewise
  LineSeg sx, sy;             // segmentation of the x-lines (length nx) / y-lines (length ny)
  DevBuf<cplx<T>> c1d, mass, gx, gy;
  DevBuf<c128> eps;           // level eps_r (fp64), level 0 aliases the fine operator's copy (not owned)
  DevBuf<cplx<T>> u, f, tmp;
  DevBuf<cplx<T>> rxs, rys;   // strip residual buffers: [line][i]
  DevBuf<unsigned long long> corner_slots;   // fp32: meeting point of the y-line and the x-line update of every PML corner point (k_lines2)
  DevBuf<cplx<T>> pcr_y, pcr_x;  // per (line, segment): alpha[K][SL] | gamma[K][SL] | binv[SL]
  DevBuf<cplx<T>> pw;         // prolongation weights of THIS level's points: wlx[nx] | wrx[nx] | wly[ny] | wry[ny]
  DevBuf<cplx<T>> rw;         // restriction weights to the next coarser level: RX[3*ncx] | RY[3*ncy]
  DevBuf<c128> rwd;           // same in fp64 (eps restriction at setup)
  cplx<T> mass_const{T(0), T(0)};
  Coef1D hc;                  // host copy of this level's 1-D coefficients in fp64, unscaled (multilevel Krylov builds A_l from it)
  OpView<T> view() const {
    OpView<T> v;
    v.nx = nx; v.ny = ny;
    v.cxm = c1d.p; v.cxp = c1d.p + nx; v.cym = c1d.p + 2 * nx; v.cyp = c1d.p + 2 * nx + ny;
    v.mass = mass.p; v.gx = gx.p; v.gy = gy.p; v.mass_const = mass_const;
    return v;
  }
};

struct MGParams {
  int cycle = FDFD_CYCLE_W, wdepth = 4, nu1 = 1, nu2 = 1, coarse_sweeps = 4;
  double beta = 0.5, wjac = 0.7, wline = 0.6, shift_growth = 0.0;
  int min_n = 2, pad = 1, max_levels = 32;
  double kh_stop = 2.0;   // measured at 4096^2 (multilevel solver, F cycles): 7 levels (kh_stop 4) 3.5 s, 6 levels (2) 2.3 s, same 105 outer
                          // iterations; BiCGSTAB 4.5 -> 4.2 s; 5 levels (1) no longer converges
};

std::vector<std::pair<int64_t, int64_t>> mg_level_sizes(const fdfd_grid_t& g, double omega, const MGParams& prm, int force_levels);

template <typename T> struct Multigrid {
  fdfd_ctx* ctx = nullptr;
  bool te = false;
  MGParams prm;
  std::vector<MGLevel<T>> lv;
  DevBuf<c128> pcr_scratch;   // setup-only scratch
  DevBuf<cplx<T>> spare;         // third level-0 buffer: lets the caller keep one result across the next apply
  double rhs_scale = 1.0;        // M is stored scaled by this factor (keeps fp32 in range); callers scale the rhs they write
  // optional device flag: kernels early-exit when it is set.  The solvers leave it NULL: the cycle only writes its own buffers, so
  // iterations enqueued past convergence are harmless, while the flag costs every (latency-bound) coarse kernel one dependent
  // global load before its first useful instruction; the Krylov update / apply kernels, which own x and r, do check it.
  const int* done = nullptr;

  // build hierarchy for the operator `op` (fine eps_r resident in op.eps)
  // first_level > 0 builds only levels >= first_level (the agglomerated coarse part of a slab-sharded solve): `op` then
  // needs its host-side members only and eps_first is the level-first_level eps_r resident in HBM
  int setup(fdfd_ctx* ctx, const FineOp& op, const MGParams& prm, int first_level = 0, const c128* eps_first = nullptr);
  int first = 0;
  int wbase = 0;                 // level the W recursion depth is counted from (a cycle started on level s by the multilevel solver sets s)
  // u0 = approx M^-1 f0 where f0 = lv[0].f (already filled).  Result pointer returned in *out (lv[0].u or .tmp)
  int apply(const cplx<T>** out);
  cplx<T>* rhs() { return lv[0].f.p; }
  int levels() const { return (int)lv.size(); }

  // building blocks (public: the slab driver runs them in lock step over several slabs with halo exchanges between)
  int cycle(int l, bool zero, int kind);
  int smooth(int l, bool zero, bool prolong = false);
  int restrict_residual(int l);
};
