// device_ops.cuh -- operator views shared by assembly, stencil, Krylov and multigrid code.
#pragma once
#include "common.cuh"

// Every level of every operator on the path has this shape:
//   (A u)[ix,iy] = W (u[ix-1]-u) + E (u[ix+1]-u) + S (u[iy-1]-u) + Nn (u[iy+1]-u) + m u        (periodic)
//   TM: W=cxm[ix] E=cxp[ix] S=cym[iy] Nn=cyp[iy]  m=mass[ix,iy]
//   TE: W=cxm[ix] gx[ix,iy]  E=cxp[ix] gx[ix+1,iy]  S=cym[iy] gy[ix,iy]  Nn=cyp[iy] gy[ix,iy+1]  m=mass_const
template <typename T> struct OpView {
  int64_t nx, ny;
  const cplx<T>* cxm; const cplx<T>* cxp; const cplx<T>* cym; const cplx<T>* cyp;
  const cplx<T>* mass;            // 2-D, TM
  const cplx<T>* gx; const cplx<T>* gy;  // 2-D, TE
  cplx<T> mass_const;
};

// a y-slab of a global grid: owned global rows [y0, y0+nyl), stored with H halo rows on each side
struct SlabInfo {
  bool on = false;
  fdfd_grid_t gg{};        // the global grid
  int64_t y0 = 0, nyl = 0; // owned rows (fine level)
  int64_t H = 0;           // fine-level halo width, a multiple of 2^(nlevels-1); level l keeps H >> l halo rows
  int nlevels = 0;         // multigrid depth, decided on the global grid
  int64_t yoff() const { return y0 - H; }  // global row of local row 0
};

// fp64 fine-grid operator resident in HBM
struct FineOp {
  fdfd_grid_t g{};         // grid the arrays are sized for (slab: Nx x (nyl + 2H), same cell size as the global grid)
  SlabInfo slab;
  int pol = FDFD_TM, ordering = FDFD_ORDER_FB;
  double omega = 0;      // frequency of the mass term
  double omega_pml = 0;  // frequency the PML s-factors are evaluated at (== omega except modulation.jl:79 sharedpml)
  DevBuf<c128> c1d;    // cxm | cxp | cym | cyp
  DevBuf<c128> mass;   // TM: w^2 eps0 L0 eps_r
  DevBuf<c128> gx, gy; // TE: 1 / grid_average(eps0 L0 eps_r, x|y)
  DevBuf<c128> eps;    // eps_r as given (kept for multigrid setup)
  c128 mass_const{0.0, 0.0};
  Coef1D hc;           // host copy of the 1-D coefficients

  int build(fdfd_ctx* ctx, const fdfd_grid_t& g, int pol, int ordering, double omega, const fdfd_c128* eps_r_any,
            double omega_pml = 0.0);
  // slab of the global grid gg: eps_local_any holds the (nyl + 2H) x Nx local rows (halo rows included, periodic wrap)
  int build_slab(fdfd_ctx* ctx, const fdfd_grid_t& gg, int ordering, double omega, const fdfd_c128* eps_local_any,
                 int64_t y0, int64_t nyl, int nlevels, int64_t halo, double omega_pml = 0.0);
  // rediscretised operator of a multigrid level (multilevel Krylov): 1-D coefficients hc_ (nx | ny entries, unscaled);
  // TM: mass = w^2 eps0 L0 eps_l, TE: inverse averaged eps_l + the constant w^2 mu0 L0 term; eps_l resident in HBM.  Only Nx, Ny of the grid member are meaningful afterwards.
  int build_level(fdfd_ctx* ctx, const fdfd_grid_t& gfine, int pol, int64_t nx, int64_t ny, const Coef1D& hc_, double omega, const c128* eps_dev);
  OpView<double> view() const {
    OpView<double> v;
    v.nx = g.Nx; v.ny = g.Ny;
    v.cxm = c1d.p; v.cxp = c1d.p + g.Nx; v.cym = c1d.p + 2 * g.Nx; v.cyp = c1d.p + 2 * g.Nx + g.Ny;
    v.mass = mass.p; v.gx = gx.p; v.gy = gy.p; v.mass_const = mass_const;
    return v;
  }
};

// ---- launch wrappers implemented in stencil.cu ------------------------------------------------
// y = A x.  TI = element type of x (c128 or c64), y is c128.  Optional fused dots (deterministic two-stage):
//   ndot = 0: none;  1: partial[0] = <d0, y>;  2: partial[0] = <y, d0>, partial[1] = <y, y>   (conjugate-linear in 1st arg)
// sideband coupling of the modulated operator (modulation.jl:95-101): y += hw * (conj(deps) x_{j+1} + deps x_{j-1})
struct Coupling { const void* xm1 = nullptr; const void* xp1 = nullptr; const c128* deps = nullptr; double hw = 0.0; };
struct DotSpec {
  int ndot = 0; const c128* d0 = nullptr; c128* partials = nullptr; int* nblocks_out = nullptr; const int* done = nullptr;
  int64_t row_lo = 0, row_hi = -1;  // slab mode: only rows [row_lo,row_hi) are computed (halo rows of y are left untouched)
};
int launch_apply(fdfd_ctx* ctx, const OpView<double>& op, bool te, const void* x, bool x_is_f32, c128* y, const DotSpec& ds,
                 const Coupling* cpl = nullptr);

// y_b = A x_b for nrhs right-hand sides sharing the operator (vectors `stride` elements apart): coefficients read once per point
int launch_apply_batched(fdfd_ctx* ctx, const OpView<double>& op, bool te, const c128* x, c128* y, int nrhs, int64_t stride);

// H/E recovery written straight into the (Nx,Ny,3) output (K9).  mode: see stencil.cu
int launch_recover(fdfd_ctx* ctx, const FineOp& op, const c128* u, int forward, std::complex<double> omega_field,
                   int te_swap, c128* fields3);
