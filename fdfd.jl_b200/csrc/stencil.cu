// stencil.cu -- K4/K5 matrix-free complex128 Yee-stencil apply, K9 field recovery, operator setup.
//
// The apply kernel is the headline HBM-bound kernel: algorithmic traffic 48 B/point (TM: read x 16,
// read mass 16, write y 16; 1-D PML coefficient arrays are L1/L2 resident).  Layout: x fastest, one thread
// per x position (128-bit loads, a warp covers 512 contiguous bytes), each thread marches ROWS rows in y
// and keeps the y-neighbours in registers, so every x value is fetched from L2/HBM once per row-block;
// x+-1 neighbours come from the same 128 B lines through L1.
#include "device_ops.cuh"
#include "reduce.cuh"

namespace {

__global__ void k_setup_tm(int64_t N, double w2eps0, const c128* __restrict__ eps, c128* __restrict__ mass) {
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    c128 e = eps[n];
    mass[n] = c128(w2eps0 * e.x, w2eps0 * e.y);
  }
}

// gx = 1 ./ grid_average(eps0*eps_r, x), gy likewise (grid.jl:157-162, driven.jl:22-23)
__global__ void k_setup_te(int64_t Nx, int64_t Ny, double eps0, const c128* __restrict__ eps, c128* __restrict__ gx,
                           c128* __restrict__ gy) {
  const int64_t N = Nx * Ny;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ix = n % Nx, iy = n / Nx;
    const int64_t ixm = ix == 0 ? Nx - 1 : ix - 1, iym = iy == 0 ? Ny - 1 : iy - 1;
    c128 e = eps[n], ew = eps[ixm + Nx * iy], es = eps[ix + Nx * iym];
    c128 a = c128(eps0 * e.x, eps0 * e.y);
    c128 aw = c128(eps0 * ew.x, eps0 * ew.y), as = c128(eps0 * es.x, eps0 * es.y);
    c128 avx = c128((a.x + aw.x) / 2, (a.y + aw.y) / 2);
    c128 avy = c128((a.x + as.x) / 2, (a.y + as.y) / 2);
    gx[n] = crecip(avx);
    gy[n] = crecip(avy);
  }
}

template <typename TI> __device__ __forceinline__ c128 ldx(const TI* p, int64_t i) { return c128(p[i]); }

constexpr int kApplyThreads = 128;

// streaming (evict-first) access for operands touched exactly once: the w^2 eps term and the output
__device__ __forceinline__ c128 ld_stream(const c128* p) { const double2 v = __ldcs(reinterpret_cast<const double2*>(p)); return c128(v.x, v.y); }
__device__ __forceinline__ void st_stream(c128* p, c128 v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }

template <typename TI, bool TE, int NDOT, int ROWS, int MINB, bool HINT>
__global__ void __launch_bounds__(kApplyThreads, MINB)
k_apply(OpView<double> op, const TI* __restrict__ x, c128* __restrict__ y, const c128* __restrict__ d0,
        c128* __restrict__ partials, const int* __restrict__ done, const TI* __restrict__ xm1, const TI* __restrict__ xp1,
        const c128* __restrict__ deps, double hw, int64_t row_lo, int64_t row_hi) {
  if (done && *done) return;
  const int64_t Nx = op.nx, Ny = op.ny;
  const int64_t ix = blockIdx.x * (int64_t)kApplyThreads + threadIdx.x;
  const int64_t iy0 = row_lo + blockIdx.y * (int64_t)ROWS;   // a slab launches over its owned rows [row_lo, row_hi) only
  double acc[NDOT > 0 ? 2 * NDOT : 1];
#pragma unroll
  for (int k = 0; k < (NDOT > 0 ? 2 * NDOT : 1); ++k) acc[k] = 0.0;
  if (ix < Nx) {
    const int64_t ixm = ix == 0 ? Nx - 1 : ix - 1, ixp = ix + 1 == Nx ? 0 : ix + 1;
    const c128 cw = op.cxm[ix], ce = op.cxp[ix];
    int64_t iym = iy0 == 0 ? Ny - 1 : iy0 - 1;
    c128 us = ldx(x, ix + Nx * iym);
    c128 uc = ldx(x, ix + Nx * iy0);
    c128 gyc = TE ? op.gy[ix + Nx * iy0] : c128(1.0, 0.0);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int64_t iy = iy0 + r;
      if (iy >= row_hi) break;
      const int64_t iyp = iy + 1 == Ny ? 0 : iy + 1;
      const int64_t n = ix + Nx * iy;
      const c128 un = ldx(x, ix + Nx * iyp);
      const c128 uw = ldx(x, ixm + Nx * iy), ue = ldx(x, ixp + Nx * iy);
      c128 W = cw, E = ce, S = op.cym[iy], Nn = op.cyp[iy], m;
      if (TE) {
        const c128 gyn = op.gy[ix + Nx * iyp];
        W = W * op.gx[n]; E = E * op.gx[ixp + Nx * iy];
        S = S * gyc; Nn = Nn * gyn;
        gyc = gyn;
        m = op.mass_const;
      } else {
        m = HINT ? ld_stream(op.mass + n) : op.mass[n];
      }
      // y = W uw + E ue + S us + N un + ((-(W+E) - (S+N)) + m) uc   (same association as the assembled matrix)
      const c128 C = ((-W - E) + (-S - Nn)) + m;
      c128 out = C * uc;
      cfma(out, W, uw); cfma(out, E, ue); cfma(out, S, us); cfma(out, Nn, un);
      if (deps) {  // sideband coupling (block-coupled MF-FDFD operator)
        const c128 de = deps[n];
        c128 cp(0.0, 0.0);
        if (xp1) cp += conj(de) * ldx(xp1, n);
        if (xm1) cp += de * ldx(xm1, n);
        out += c128(hw * cp.x, hw * cp.y);
      }
      if (HINT) st_stream(y + n, out); else y[n] = out;
      if constexpr (NDOT == 1) {  // <d0, y> = sum conj(d0) y
        const c128 d = d0[n];
        const c128 p = cmulc(d, out);
        acc[0] += p.x; acc[1] += p.y;
      } else if constexpr (NDOT == 2) {  // <y, d0>, <y, y>
        const c128 d = d0[n];
        const c128 p = cmulc(out, d);
        acc[0] += p.x; acc[1] += p.y;
        acc[2] += norm2(out);
      }
      us = uc; uc = un;
    }
  }
  if constexpr (NDOT > 0) {
    const int64_t b = blockIdx.y * (int64_t)gridDim.x + blockIdx.x;
    block_reduce_store<kApplyThreads, 2 * NDOT>(acc, reinterpret_cast<double*>(partials) + b * 2 * NDOT);
  }
}

// K9: comp0 = u, comp1 = k1 * D1 u, comp2 = k2 * D2 u with D = stretched backward or forward differences.
// TM: comp1 = kx * Dy u (hx), comp2 = ky * Dx u (hy).     driven.jl:40-41 (backward) / modulation.jl:112-113, eigen.jl:90-91 (forward)
// TE: comp1 = k * g1 * Dyb u (ex), comp2 = k * g2 * (-Dxb u) (ey), (g1,g2) = (gy,gx) driven.jl:50-51 or (gx,gy) eigen.jl:108-109
template <bool TE>
__global__ void k_recover(int64_t Nx, int64_t Ny, const c128* __restrict__ u, const c128* __restrict__ sx,
                          const c128* __restrict__ sy, double ax, double ay, int forward, c128 k1, c128 k2,
                          const c128* __restrict__ g1, const c128* __restrict__ g2, c128* __restrict__ out) {
  const int64_t N = Nx * Ny;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ix = n % Nx, iy = n / Nx;
    const c128 uc = u[n];
    c128 dxu, dyu;
    if (forward) {
      const int64_t ixp = ix + 1 == Nx ? 0 : ix + 1, iyp = iy + 1 == Ny ? 0 : iy + 1;
      // row of S*δf: values s*(1/d)*(-1) on the diagonal and s*(1/d)*(+1) on the +1 neighbour
      const c128 sxv = sx[ix], syv = sy[iy];
      dxu = c128(sxv.x * -ax, sxv.y * -ax) * uc + c128(sxv.x * ax, sxv.y * ax) * u[ixp + Nx * iy];
      dyu = c128(syv.x * -ay, syv.y * -ay) * uc + c128(syv.x * ay, syv.y * ay) * u[ix + Nx * iyp];
    } else {
      const int64_t ixm = ix == 0 ? Nx - 1 : ix - 1, iym = iy == 0 ? Ny - 1 : iy - 1;
      const c128 sxv = sx[ix], syv = sy[iy];
      dxu = c128(sxv.x * -ax, sxv.y * -ax) * u[ixm + Nx * iy] + c128(sxv.x * ax, sxv.y * ax) * uc;
      dyu = c128(syv.x * -ay, syv.y * -ay) * u[ix + Nx * iym] + c128(syv.x * ay, syv.y * ay) * uc;
    }
    c128 c1, c2;
    if (TE) { c1 = k1 * (g1[n] * dyu); c2 = k2 * (g2[n] * (-dxu)); }
    else    { c1 = k1 * dyu; c2 = k2 * dxu; }
    out[n] = uc; out[N + n] = c1; out[2 * N + n] = c2;
  }
}

}  // namespace

int FineOp::build(fdfd_ctx* ctx, const fdfd_grid_t& g_, int pol_, int ordering_, double omega_, const fdfd_c128* eps_r_any,
                  double omega_pml_) {
  g = g_; pol = pol_; ordering = ordering_; omega = omega_; omega_pml = omega_pml_ > 0 ? omega_pml_ : omega_;
  const int64_t N = g.Nx * g.Ny;
  ARG_CHECK(ctx, !(pol == FDFD_TE && ordering != FDFD_ORDER_FB), "TE is defined for the f.b ordering only");
  const double eps0 = kEps0 * g.L0, mu0 = kMu0 * g.L0;
  // TM: mu0^-1 folded into the 1-D coefficients (driven.jl:35 `δxf*μ₀^-1*δxb`); TE: none (driven.jl:45)
  host_coef_fine(g, omega_pml, ordering, pol == FDFD_TM ? 1.0 / mu0 : 1.0, hc);
  CUDA_TRY(ctx, c1d.alloc(2 * g.Nx + 2 * g.Ny));
  std::vector<std::complex<double>> pack;
  pack.insert(pack.end(), hc.cxm.begin(), hc.cxm.end()); pack.insert(pack.end(), hc.cxp.begin(), hc.cxp.end());
  pack.insert(pack.end(), hc.cym.begin(), hc.cym.end()); pack.insert(pack.end(), hc.cyp.begin(), hc.cyp.end());
  CUDA_TRY(ctx, cudaMemcpyAsync(c1d.p, pack.data(), pack.size() * sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // pack is a local
  CUDA_TRY(ctx, eps.alloc(N));
  FDFD_TRY(fdfd_copy_in(ctx, eps.p, eps_r_any, N * sizeof(c128)));
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((N + threads - 1) / threads, (int64_t)ctx->num_sms * 16);
  if (pol == FDFD_TM) {
    CUDA_TRY(ctx, mass.alloc(N));
    // ω^2*Tϵ with Tϵ = ϵ₀*ϵᵣ (driven.jl:21,35): (ω^2) * (eps0*eps) -- fold the two real factors
    k_setup_tm<<<blocks, threads, 0, ctx->stream>>>(N, omega * omega * eps0, eps.p, mass.p);
    KLAUNCH(ctx);
  } else {
    CUDA_TRY(ctx, gx.alloc(N)); CUDA_TRY(ctx, gy.alloc(N));
    k_setup_te<<<blocks, threads, 0, ctx->stream>>>(g.Nx, g.Ny, eps0, eps.p, gx.p, gy.p);
    KLAUNCH(ctx);
    mass_const = c128(omega * omega * mu0, 0.0);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return FDFD_OK;
}

// tuning variants (FDFD_APPLY_VARIANT): rows marched per thread, min CTAs/SM (register cap), streaming hints
static int apply_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FDFD_APPLY_VARIANT"); v = e ? atoi(e) : 0; if (v < 0 || v > 7) v = 0; }
  return v;
}
static int variant_rows(int v) { static const int rows[8] = {4, 8, 16, 4, 8, 16, 32, 8}; return rows[v]; }

int FineOp::build_slab(fdfd_ctx* ctx, const fdfd_grid_t& gg, int ordering_, double omega_, const fdfd_c128* eps_local_any,
                       int64_t y0, int64_t nyl, int nlevels, int64_t halo, double omega_pml_) {
  pol = FDFD_TM; ordering = ordering_; omega = omega_; omega_pml = omega_pml_ > 0 ? omega_pml_ : omega_;
  slab.on = true; slab.gg = gg; slab.y0 = y0; slab.nyl = nyl; slab.nlevels = nlevels; slab.H = halo;
  ARG_CHECK(ctx, halo >= 1 && halo % ((int64_t)1 << (nlevels - 1)) == 0 && halo <= nyl, "slab halo must be a multiple of 2^(levels-1) and at most the slab height");
  ARG_CHECK(ctx, nyl % ((int64_t)1 << (nlevels - 1)) == 0 && y0 % ((int64_t)1 << (nlevels - 1)) == 0 && gg.Ny % ((int64_t)1 << (nlevels - 1)) == 0,
            "slab rows must be divisible by 2^(levels-1)");
  const int64_t nloc = nyl + 2 * slab.H;
  g = gg; g.Ny = nloc; g.Npml_y = 0;
  const double dyg = grid_dy(gg);
  g.y0 = gg.y0 + dyg * (double)slab.yoff(); g.y1 = g.y0 + dyg * (double)nloc;
  const int64_t N = g.Nx * nloc;
  const double eps0 = kEps0 * gg.L0, mu0 = kMu0 * gg.L0;
  Coef1D gh;
  host_coef_fine(gg, omega_pml, ordering, 1.0 / mu0, gh);
  hc.cxm = gh.cxm; hc.cxp = gh.cxp; hc.cym.resize(nloc); hc.cyp.resize(nloc);
  for (int64_t i = 0; i < nloc; ++i) {
    const int64_t gi = ((slab.yoff() + i) % gg.Ny + gg.Ny) % gg.Ny;
    hc.cym[i] = gh.cym[gi]; hc.cyp[i] = gh.cyp[gi];
  }
  CUDA_TRY(ctx, c1d.alloc(2 * g.Nx + 2 * nloc));
  std::vector<std::complex<double>> pack;
  pack.insert(pack.end(), hc.cxm.begin(), hc.cxm.end()); pack.insert(pack.end(), hc.cxp.begin(), hc.cxp.end());
  pack.insert(pack.end(), hc.cym.begin(), hc.cym.end()); pack.insert(pack.end(), hc.cyp.begin(), hc.cyp.end());
  CUDA_TRY(ctx, cudaMemcpyAsync(c1d.p, pack.data(), pack.size() * sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(ctx, eps.alloc(N));
  FDFD_TRY(fdfd_copy_in(ctx, eps.p, eps_local_any, N * sizeof(c128)));
  CUDA_TRY(ctx, mass.alloc(N));
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((N + threads - 1) / threads, (int64_t)ctx->num_sms * 16);
  k_setup_tm<<<blocks, threads, 0, ctx->stream>>>(N, omega * omega * eps0, eps.p, mass.p);
  KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  return FDFD_OK;
}

int FineOp::build_level(fdfd_ctx* ctx, const fdfd_grid_t& gfine, int pol_, int64_t nx, int64_t ny, const Coef1D& hc_, double omega_, const c128* eps_dev) {
  g = gfine; g.Nx = nx; g.Ny = ny;
  pol = pol_; ordering = FDFD_ORDER_FB; omega = omega_; omega_pml = omega_;
  hc = hc_;
  ARG_CHECK(ctx, (int64_t)hc.cxm.size() == nx && (int64_t)hc.cxp.size() == nx && (int64_t)hc.cym.size() == ny && (int64_t)hc.cyp.size() == ny,
            "internal: level coefficient arrays do not match the level size");
  const int64_t N = nx * ny;
  const double eps0 = kEps0 * g.L0, mu0 = kMu0 * g.L0;
  CUDA_TRY(ctx, c1d.alloc(2 * nx + 2 * ny));
  std::vector<std::complex<double>> pack;
  pack.insert(pack.end(), hc.cxm.begin(), hc.cxm.end()); pack.insert(pack.end(), hc.cxp.begin(), hc.cxp.end());
  pack.insert(pack.end(), hc.cym.begin(), hc.cym.end()); pack.insert(pack.end(), hc.cyp.begin(), hc.cyp.end());
  CUDA_TRY(ctx, cudaMemcpyAsync(c1d.p, pack.data(), pack.size() * sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // pack is a local
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((N + threads - 1) / threads, (int64_t)ctx->num_sms * 16);
  if (pol == FDFD_TM) {
    CUDA_TRY(ctx, mass.alloc(N));
    k_setup_tm<<<blocks, threads, 0, ctx->stream>>>(N, omega * omega * eps0, eps_dev, mass.p);
  } else {
    CUDA_TRY(ctx, gx.alloc(N)); CUDA_TRY(ctx, gy.alloc(N));
    k_setup_te<<<blocks, threads, 0, ctx->stream>>>(nx, ny, eps0, eps_dev, gx.p, gy.p);
    mass_const = c128(omega * omega * mu0, 0.0);
  }
  KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  return FDFD_OK;
}


// ---- B right-hand sides sharing ONE operator (the batch axis of the reference's sweep, driven.jl:11, restricted to sources that share
// a frequency): the coefficients -- above all the 16 B/pt of w^2 eps (TM) or the two inverse averaged eps arrays (TE) -- are read once
// per point and applied to B vectors: (32 B + 16) / B algorithmic bytes per point and right-hand side instead of 48.
// Same marching scheme and the same arithmetic as k_apply (the compiler contracts the FMAs differently: results agree to rounding).
template <bool TE, int B, int ROWS>
__global__ void __launch_bounds__(kApplyThreads, B >= 4 ? 3 : 6)
k_apply_batched(OpView<double> op, const c128* __restrict__ x, c128* __restrict__ y, int64_t stride) {
  const int64_t Nx = op.nx, Ny = op.ny;
  const int64_t ix = blockIdx.x * (int64_t)kApplyThreads + threadIdx.x;
  const int64_t iy0 = blockIdx.y * (int64_t)ROWS;
  if (ix >= Nx) return;
  const int64_t ixm = ix == 0 ? Nx - 1 : ix - 1, ixp = ix + 1 == Nx ? 0 : ix + 1;
  const c128 cw = op.cxm[ix], ce = op.cxp[ix];
  const int64_t iym0 = iy0 == 0 ? Ny - 1 : iy0 - 1;
  c128 us[B], uc[B];
#pragma unroll
  for (int b = 0; b < B; ++b) { us[b] = x[b * stride + ix + Nx * iym0]; uc[b] = x[b * stride + ix + Nx * iy0]; }
  c128 gyc = TE ? op.gy[ix + Nx * iy0] : c128(1.0, 0.0);
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int64_t iy = iy0 + r;
    if (iy >= Ny) break;
    const int64_t iyp = iy + 1 == Ny ? 0 : iy + 1;
    const int64_t n = ix + Nx * iy;
    c128 W = cw, E = ce, S = op.cym[iy], Nn = op.cyp[iy], m;
    if (TE) {
      const c128 gyn = op.gy[ix + Nx * iyp];
      W = W * op.gx[n]; E = E * op.gx[ixp + Nx * iy];
      S = S * gyc; Nn = Nn * gyn;
      gyc = gyn;
      m = op.mass_const;
    } else {
      m = ld_stream(op.mass + n);
    }
    const c128 C = ((-W - E) + (-S - Nn)) + m;
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const c128* xb = x + b * stride;
      const c128 un = xb[ix + Nx * iyp];
      const c128 uw = xb[ixm + Nx * iy], ue = xb[ixp + Nx * iy];
      c128 out = C * uc[b];
      cfma(out, W, uw); cfma(out, E, ue); cfma(out, S, us[b]); cfma(out, Nn, un);
      st_stream(y + b * stride + n, out);
      us[b] = uc[b]; uc[b] = un;
    }
  }
}

template <bool TE, int B>
static int launch_apply_batched_b(fdfd_ctx* ctx, const OpView<double>& op, const c128* x, c128* y, int64_t stride) {
  constexpr int ROWS = 4;
  dim3 grid((unsigned)((op.nx + kApplyThreads - 1) / kApplyThreads), (unsigned)((op.ny + ROWS - 1) / ROWS));
  k_apply_batched<TE, B, ROWS><<<grid, kApplyThreads, 0, ctx->stream>>>(op, x, y, stride);
  KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  return FDFD_OK;
}

// y_b = A x_b, b < nrhs; vectors stride elements apart.  Any nrhs: chunks of 8 / 4 / 2 / 1.
int launch_apply_batched(fdfd_ctx* ctx, const OpView<double>& op, bool te, const c128* x, c128* y, int nrhs, int64_t stride) {
  int done = 0;
  while (done < nrhs) {
    const int left = nrhs - done;
    const int B = left >= 8 ? 8 : left >= 4 ? 4 : left >= 2 ? 2 : 1;
    const c128* xb = x + (size_t)done * stride; c128* yb = y + (size_t)done * stride;
#define GO(BV) (te ? launch_apply_batched_b<true, BV>(ctx, op, xb, yb, stride) : launch_apply_batched_b<false, BV>(ctx, op, xb, yb, stride))
    FDFD_TRY(B == 8 ? GO(8) : B == 4 ? GO(4) : B == 2 ? GO(2) : GO(1));
#undef GO
    done += B;
  }
  return FDFD_OK;
}

template <typename TI, bool TE, int NDOT, int ROWS, int MINB, bool HINT>
static int launch_apply_v(fdfd_ctx* ctx, const OpView<double>& op, const TI* x, c128* y, const DotSpec& ds, const Coupling* cpl) {
  const int64_t row_hi = ds.row_hi < 0 ? op.ny : ds.row_hi;
  dim3 grid((unsigned)((op.nx + kApplyThreads - 1) / kApplyThreads), (unsigned)((row_hi - ds.row_lo + ROWS - 1) / ROWS));
  if (ds.nblocks_out) *ds.nblocks_out = (int)(grid.x * grid.y);
  k_apply<TI, TE, NDOT, ROWS, MINB, HINT><<<grid, kApplyThreads, 0, ctx->stream>>>(op, x, y, ds.d0, ds.partials, ds.done,
      cpl ? (const TI*)cpl->xm1 : nullptr, cpl ? (const TI*)cpl->xp1 : nullptr, cpl ? cpl->deps : nullptr, cpl ? cpl->hw : 0.0,
      ds.row_lo, row_hi);
  KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  return FDFD_OK;
}

template <typename TI, bool TE, int NDOT>
static int launch_apply_t(fdfd_ctx* ctx, const OpView<double>& op, const TI* x, c128* y, const DotSpec& ds, const Coupling* cpl) {
  switch (apply_variant()) {
    case 1: return launch_apply_v<TI, TE, NDOT, 8, 8, true>(ctx, op, x, y, ds, cpl);
    case 2: return launch_apply_v<TI, TE, NDOT, 16, 8, true>(ctx, op, x, y, ds, cpl);
    case 3: return launch_apply_v<TI, TE, NDOT, 4, 8, true>(ctx, op, x, y, ds, cpl);
    case 4: return launch_apply_v<TI, TE, NDOT, 8, 1, true>(ctx, op, x, y, ds, cpl);
    case 5: return launch_apply_v<TI, TE, NDOT, 16, 6, true>(ctx, op, x, y, ds, cpl);
    case 6: return launch_apply_v<TI, TE, NDOT, 32, 8, true>(ctx, op, x, y, ds, cpl);
    case 7: return launch_apply_v<TI, TE, NDOT, 8, 1, false>(ctx, op, x, y, ds, cpl);
    default: return launch_apply_v<TI, TE, NDOT, 4, 8, true>(ctx, op, x, y, ds, cpl);
  }
}

int apply_num_blocks(int64_t nx, int64_t ny) {
  const int rows = variant_rows(apply_variant());
  return (int)(((nx + kApplyThreads - 1) / kApplyThreads) * ((ny + rows - 1) / rows));
}

int launch_apply(fdfd_ctx* ctx, const OpView<double>& op, bool te, const void* x, bool x_is_f32, c128* y, const DotSpec& ds,
                 const Coupling* cpl) {
#define DISPATCH(TI, TEV)                                                                        \
  switch (ds.ndot) {                                                                             \
    case 0: return launch_apply_t<TI, TEV, 0>(ctx, op, (const TI*)x, y, ds, cpl);                     \
    case 1: return launch_apply_t<TI, TEV, 1>(ctx, op, (const TI*)x, y, ds, cpl);                     \
    default: return launch_apply_t<TI, TEV, 2>(ctx, op, (const TI*)x, y, ds, cpl);                    \
  }
  if (x_is_f32) { if (te) { DISPATCH(c64, true) } else { DISPATCH(c64, false) } }
  else          { if (te) { DISPATCH(c128, true) } else { DISPATCH(c128, false) } }
#undef DISPATCH
  return FDFD_OK;
}

int launch_recover(fdfd_ctx* ctx, const FineOp& op, const c128* u, int forward, std::complex<double> omega_field,
                   int te_swap, c128* fields3) {
  const fdfd_grid_t& g = op.g;
  const int64_t N = g.Nx * g.Ny;
  const double mu0 = kMu0 * g.L0;
  // inverse s-factors of the requested difference direction, frozen at the operator's omega
  std::vector<std::complex<double>> sx, sy;
  if (op.slab.on) {  // s-factors of the GLOBAL grid, y-array sliced to the slab's local rows
    std::vector<std::complex<double>> syg;
    host_sfactor(op.slab.gg, 0, forward, op.omega_pml, sx); host_sfactor(op.slab.gg, 1, forward, op.omega_pml, syg);
    sy.resize(g.Ny);
    for (int64_t i = 0; i < g.Ny; ++i) sy[i] = syg[((op.slab.yoff() + i) % op.slab.gg.Ny + op.slab.gg.Ny) % op.slab.gg.Ny];
  } else {
    host_sfactor(g, 0, forward, op.omega_pml, sx); host_sfactor(g, 1, forward, op.omega_pml, sy);
  }
  for (auto& z : sx) z = 1.0 / z;
  for (auto& z : sy) z = 1.0 / z;
  DevBuf<c128> ds;
  CUDA_TRY(ctx, ds.alloc(g.Nx + g.Ny));
  CUDA_TRY(ctx, cudaMemcpyAsync(ds.p, sx.data(), g.Nx * sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(ds.p + g.Nx, sy.data(), g.Ny * sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
  const std::complex<double> I(0.0, 1.0);
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((N + threads - 1) / threads, (int64_t)ctx->num_sms * 16);
  if (op.pol == FDFD_TM) {
    // hx = -1/1im/ω/μ₀ * Dy ez ; hy = 1/1im/ω/μ₀ * Dx ez   (left-to-right like Julia)
    const std::complex<double> k1 = ((-1.0 / I) / omega_field) / mu0, k2 = ((1.0 / I) / omega_field) / mu0;
    k_recover<false><<<blocks, threads, 0, ctx->stream>>>(g.Nx, g.Ny, u, ds.p, ds.p + g.Nx, 1.0 / grid_dx(g), 1.0 / grid_dy(g),
                                                         forward, to_c128(k1), to_c128(k2), nullptr, nullptr, fields3);
  } else {
    // ex = 1/1im/ω * T1 * Dyb hz ; ey = 1/1im/ω * T2 * (-Dxb hz)
    const std::complex<double> k = (1.0 / I) / omega_field;
    const c128* g1 = te_swap ? op.gx.p : op.gy.p;
    const c128* g2 = te_swap ? op.gy.p : op.gx.p;
    k_recover<true><<<blocks, threads, 0, ctx->stream>>>(g.Nx, g.Ny, u, ds.p, ds.p + g.Nx, 1.0 / grid_dx(g), 1.0 / grid_dy(g),
                                                        forward, to_c128(k), to_c128(k), g1, g2, fields3);
  }
  KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // ds, sx, sy are locals
  return FDFD_OK;
}

extern "C" int fdfd_apply_operator(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, int ordering, double omega,
                                   const fdfd_c128* eps_r, const fdfd_c128* x, fdfd_c128* y) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, pol == FDFD_TM || pol == FDFD_TE, "pol must be FDFD_TM or FDFD_TE");
  ARG_CHECK(ctx, ordering == FDFD_ORDER_FB || ordering == FDFD_ORDER_BF, "bad ordering");
  ARG_CHECK(ctx, eps_r && x && y, "NULL argument");
  ARG_CHECK(ctx, omega > 0, "omega must be > 0");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int64_t N = g->Nx * g->Ny;
  FineOp op;
  FDFD_TRY(op.build(ctx, *g, pol, ordering, omega, eps_r));
  DevBuf<c128> dx, dy;
  CUDA_TRY(ctx, dx.alloc(N)); CUDA_TRY(ctx, dy.alloc(N));
  FDFD_TRY(fdfd_copy_in(ctx, dx.p, x, N * sizeof(c128)));
  DotSpec ds;
  FDFD_TRY(launch_apply(ctx, op.view(), pol == FDFD_TE, dx.p, false, dy.p, ds));
  FDFD_TRY(fdfd_copy_out(ctx, y, dy.p, N * sizeof(c128)));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FDFD_OK;
}

// apply_operator for nrhs right-hand sides that share the operator: x, y are nrhs x (Nx,Ny), column-major, one after the other
extern "C" int fdfd_apply_operator_batched(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, int ordering, double omega,
                                           const fdfd_c128* eps_r, int nrhs, const fdfd_c128* x, fdfd_c128* y) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, pol == FDFD_TM || pol == FDFD_TE, "pol must be FDFD_TM or FDFD_TE");
  ARG_CHECK(ctx, ordering == FDFD_ORDER_FB || ordering == FDFD_ORDER_BF, "bad ordering");
  ARG_CHECK(ctx, eps_r && x && y, "NULL argument");
  ARG_CHECK(ctx, omega > 0, "omega must be > 0");
  ARG_CHECK(ctx, nrhs >= 1 && nrhs <= 64, "nrhs must be in [1, 64]");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int64_t N = g->Nx * g->Ny;
  FineOp op;
  FDFD_TRY(op.build(ctx, *g, pol, ordering, omega, eps_r));
  DevBuf<c128> dx, dy;
  CUDA_TRY(ctx, dx.alloc((size_t)nrhs * N)); CUDA_TRY(ctx, dy.alloc((size_t)nrhs * N));
  FDFD_TRY(fdfd_copy_in(ctx, dx.p, x, (size_t)nrhs * N * sizeof(c128)));
  FDFD_TRY(launch_apply_batched(ctx, op.view(), pol == FDFD_TE, dx.p, dy.p, nrhs, N));
  FDFD_TRY(fdfd_copy_out(ctx, y, dy.p, (size_t)nrhs * N * sizeof(c128)));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FDFD_OK;
}
