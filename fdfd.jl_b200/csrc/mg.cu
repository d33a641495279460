// mg.cu -- shifted-Laplacian geometric multigrid (see mg.cuh for the design).
#include "mg.cuh"
#include <cmath>
#include <string>

namespace {

constexpr int kMgThreads = 128;
constexpr int kMgRows = 4;

// coefficients of row (ix,iy): W,E,S,N couplings and the diagonal C = m - W - E - S - N
template <typename T, bool TE>
__device__ __forceinline__ void row_coefs(const OpView<T>& op, int64_t ix, int64_t iy, int64_t ixp, int64_t iyp,
                                          cplx<T>& W, cplx<T>& E, cplx<T>& S, cplx<T>& Nn, cplx<T>& C) {
  const int64_t nx = op.nx;
  W = op.cxm[ix]; E = op.cxp[ix]; S = op.cym[iy]; Nn = op.cyp[iy];
  cplx<T> m;
  if (TE) {
    const int64_t n = ix + nx * iy;
    W = W * op.gx[n]; E = E * op.gx[ixp + nx * iy];
    S = S * op.gy[n]; Nn = Nn * op.gy[ix + nx * iyp];
    m = op.mass_const;
  } else {
    m = op.mass[ix + nx * iy];
  }
  C = m - W - E - S - Nn;
}

template <typename T, bool TE>
__device__ __forceinline__ cplx<T> residual_at(const OpView<T>& op, const cplx<T>* __restrict__ u,
                                               const cplx<T>* __restrict__ f, int64_t ix, int64_t iy, cplx<T>* Cout) {
  const int64_t nx = op.nx, ny = op.ny;
  const int64_t ixm = ix == 0 ? nx - 1 : ix - 1, ixp = ix + 1 == nx ? 0 : ix + 1;
  const int64_t iym = iy == 0 ? ny - 1 : iy - 1, iyp = iy + 1 == ny ? 0 : iy + 1;
  cplx<T> W, E, S, Nn, C;
  row_coefs<T, TE>(op, ix, iy, ixp, iyp, W, E, S, Nn, C);
  cplx<T> r = f[ix + nx * iy];
  r -= C * u[ix + nx * iy];
  r -= W * u[ixm + nx * iy]; r -= E * u[ixp + nx * iy];
  r -= S * u[ix + nx * iym]; r -= Nn * u[ix + nx * iyp];
  if (Cout) *Cout = C;
  return r;
}

__device__ __forceinline__ bool in_strip(int64_t i, int64_t n, int np) { return i < np || i >= n - np; }
__device__ __forceinline__ int64_t strip_line(int64_t i, int64_t n, int np) { return i < np ? i : i - (n - 2 * np); }
__device__ __forceinline__ int64_t strip_index(int64_t c, int64_t n, int np) { return c < np ? c : c + (n - 2 * np); }
// y-strip rows as two ranges [a0,a1) U [b0,b1) (a slab owns only part of the global PML rows, see slab.cu)
__device__ __forceinline__ bool ys_in(const YS& s, int iy) { return (iy >= s.a0 && iy < s.a1) || (iy >= s.b0 && iy < s.b1); }
__device__ __forceinline__ int ys_line(const YS& s, int iy) { return iy < s.a1 ? iy - s.a0 : (s.a1 - s.a0) + (iy - s.b0); }
__device__ __forceinline__ int ys_row(const YS& s, int c) { const int na = s.a1 - s.a0; return c < na ? s.a0 + c : s.b0 + (c - na); }

// ---- level setup --------------------------------------------------------------------------
// TM: mass = (1 - i beta) w^2 eps0 eps ;  TE: gx, gy = 1/grid_average(eps0 eps)
template <typename T, bool TE>
__global__ void k_level_setup(int64_t nx, int64_t ny, const c128* __restrict__ eps, double w2eps0, double beta, double eps0,
                              double growth_kfac, cplx<T>* __restrict__ mass, cplx<T>* __restrict__ gx, cplx<T>* __restrict__ gy) {
  const int64_t N = nx * ny;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    const c128 e = eps[n];
    if (!TE) {
      // (1 - i beta_eff) * w^2 eps0 * e ; beta_eff grows with the local k^2 h_l^2 so that coarse levels whose
      // re-discretised dispersion is wrong (kh >~ 1) become damped instead of resonant
      const double be = fmax(beta, growth_kfac * e.x);
      mass[n] = cplx<T>(T(w2eps0 * (e.x + be * e.y)), T(w2eps0 * (e.y - be * e.x)));
    } else {
      const int64_t ix = n % nx, iy = n / nx;
      const int64_t ixm = ix == 0 ? nx - 1 : ix - 1, iym = iy == 0 ? ny - 1 : iy - 1;
      const c128 ew = eps[ixm + nx * iy], es = eps[ix + nx * iym];
      const c128 ax = c128(eps0 * (e.x + ew.x) / 2, eps0 * (e.y + ew.y) / 2);
      const c128 ay = c128(eps0 * (e.x + es.x) / 2, eps0 * (e.y + es.y) / 2);
      gx[n] = cplx<T>(crecip(ax)); gy[n] = cplx<T>(crecip(ay));
    }
  }
}

// 1-D transfer stencil of coarse point I (fine 2I): fine indices of its left / centre / right contributors.  The
// weights come from per-level arrays built on the host (Galerkin-consistent in the stretched coordinate, see setup).
__device__ __forceinline__ void tr_index(int64_t I, int64_t nf, int64_t idx[3]) {
  idx[1] = 2 * I;
  idx[0] = I == 0 ? nf - 1 : 2 * I - 1;           // weight is 0 when that point is itself a coarse point (odd nf seam)
  idx[2] = 2 * I + 1 <= nf - 1 ? 2 * I + 1 : 0;   // weight is 0 when missing
}

// eps restriction with the same (volume weighted) restriction as the residual
__global__ void k_restrict_eps(int64_t nxf, int64_t nyf, int64_t nxc, int64_t nyc, const c128* __restrict__ rx,
                               const c128* __restrict__ ry, const c128* __restrict__ ef, c128* __restrict__ ec) {
  const int64_t Nc = nxc * nyc;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < Nc; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = n % nxc, J = n / nxc;
    int64_t xi[3], yi[3];
    tr_index(I, nxf, xi); tr_index(J, nyf, yi);
    c128 acc(0.0, 0.0);
    for (int b = 0; b < 3; ++b)
      for (int a = 0; a < 3; ++a) {
        const c128 w = rx[3 * I + a] * ry[3 * J + b];
        if (w.x == 0.0 && w.y == 0.0) continue;
        acc += w * ef[xi[a] + nxf * yi[b]];
      }
    ec[n] = acc;
  }
}

// ---- PCR setup: one CTA per (line, segment), fp64, global scratch (a,b,c ping-pong) -------------------------
// mode 0: y-lines (column ix = strip_index(c)), a=S, c=N.  mode 1: x-lines (row iy), a=W, c=E.  Couplings are cut at
// the segment ends (at the ends of the line these are the periodic wrap couplings: they link the two deepest PML cells
// and stay in the lagged part of the splitting).
template <typename T, bool TE>
__global__ void k_pcr_setup(OpView<T> op, int mode, int np, YS ys, LineSeg sg, c128* __restrict__ scratch, cplx<T>* __restrict__ mult) {
  const int64_t nx = op.nx, ny = op.ny;
  const int64_t c = blockIdx.x / sg.nseg;
  const int j = blockIdx.x % sg.nseg;
  const int lo = sg.lo(j), n = sg.hi(j) - lo, K = sg.K, SL = sg.SL;
  const int64_t fixed = mode == 0 ? strip_index(c, nx, np) : (int64_t)ys_row(ys, (int)c);
  c128* a0 = scratch + (size_t)blockIdx.x * 6 * SL; c128* b0 = a0 + SL; c128* c0 = b0 + SL;
  c128* a1 = c0 + SL; c128* b1 = a1 + SL; c128* c1 = b1 + SL;
  cplx<T>* alpha = mult + (size_t)blockIdx.x * (2 * K + 1) * SL;
  cplx<T>* gamma = alpha + (size_t)K * SL;
  cplx<T>* binv = gamma + (size_t)K * SL;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t ix = mode == 0 ? fixed : lo + i, iy = mode == 0 ? lo + i : fixed;
    const int64_t ixp = ix + 1 == nx ? 0 : ix + 1, iyp = iy + 1 == ny ? 0 : iy + 1;
    cplx<T> W, E, S, Nn, C;
    row_coefs<T, TE>(op, ix, iy, ixp, iyp, W, E, S, Nn, C);
    c128 l = mode == 0 ? c128(S) : c128(W), h = mode == 0 ? c128(Nn) : c128(E);
    if (i == 0) l = c128(0.0, 0.0);
    if (i == n - 1) h = c128(0.0, 0.0);
    a0[i] = l; b0[i] = c128(C); c0[i] = h;
  }
  __syncthreads();
  for (int k = 0; k < K; ++k) {
    const int s = 1 << k;
    for (int i = threadIdx.x; i < SL; i += blockDim.x) {
      c128 al(0.0, 0.0), ga(0.0, 0.0);
      if (i < n) {
        c128 bn = b0[i], an(0.0, 0.0), cn(0.0, 0.0);
        if (i >= s) { al = -cdiv(a0[i], b0[i - s]); bn += al * c0[i - s]; an = al * a0[i - s]; }
        if (i + s < n) { ga = -cdiv(c0[i], b0[i + s]); bn += ga * a0[i + s]; cn = ga * c0[i + s]; }
        a1[i] = an; b1[i] = bn; c1[i] = cn;
      }
      alpha[(size_t)k * SL + i] = cplx<T>(al); gamma[(size_t)k * SL + i] = cplx<T>(ga);
    }
    __syncthreads();
    c128* t;
    t = a0; a0 = a1; a1 = t; t = b0; b0 = b1; b1 = t; t = c0; c0 = c1; c1 = t;
  }
  for (int i = threadIdx.x; i < SL; i += blockDim.x) binv[i] = i < n ? cplx<T>(crecip(b0[i])) : cplx<T>(T(0), T(0));
}


// =====================================================================================================
// Fused smoother (2 launches per sweep instead of 4; the coarse levels are launch-latency bound):
//   k_smooth2: residual of every point from the same iterate; Jacobi update outside the strips; residual capture
//              for the strip columns AND the strip rows; optionally the coarse-grid correction is applied on the
//              fly (PROLONG): the iterate it reads is u + P u_c, evaluated per stencil point, and written back.
//   k_lines2:  all y-lines (strip columns, updating the rows outside the y-strip) and all x-lines (strip rows,
//              including the corners) in one launch.
// =====================================================================================================
template <typename T> struct ProlongView {
  int64_t nxc, nyc;
  const cplx<T>* pwx; const cplx<T>* pwy; const cplx<T>* uc;
};

template <typename T, bool PROLONG>
__device__ __forceinline__ cplx<T> iterate_at(const cplx<T>* __restrict__ u, const ProlongView<T>& pv, int64_t nx, int64_t ny,
                                              int64_t ix, int64_t iy) {
  cplx<T> v = u[ix + nx * iy];
  if (PROLONG) {
    const int64_t I0 = ix >> 1, J0 = iy >> 1;
    const bool ox = ix & 1, oy = iy & 1;
    const int64_t I1 = ox ? (I0 + 1 == pv.nxc ? 0 : I0 + 1) : I0;
    const int64_t J1 = oy ? (J0 + 1 == pv.nyc ? 0 : J0 + 1) : J0;
    const cplx<T> one(T(1), T(0)), zero(T(0), T(0));
    const cplx<T> wxl = ox ? pv.pwx[ix] : one, wxr = ox ? pv.pwx[nx + ix] : zero;
    const cplx<T> wyl = oy ? pv.pwy[iy] : one, wyr = oy ? pv.pwy[ny + iy] : zero;
    cplx<T> lo = wxl * pv.uc[I0 + pv.nxc * J0];
    if (ox) lo += wxr * pv.uc[I1 + pv.nxc * J0];
    cplx<T> c = wyl * lo;
    if (oy) {
      cplx<T> hi = wxl * pv.uc[I0 + pv.nxc * J1];
      if (ox) hi += wxr * pv.uc[I1 + pv.nxc * J1];
      c += wyr * hi;
    }
    v += c;
  }
  return v;
}

template <typename T, bool TE, bool ZERO, bool PROLONG>
__global__ void __launch_bounds__(kMgThreads)
k_smooth2(OpView<T> op, const cplx<T>* __restrict__ u, const cplx<T>* __restrict__ f, cplx<T>* __restrict__ out,
          cplx<T>* __restrict__ rxs, cplx<T>* __restrict__ rys, int npx, YS ysr, T wj, ProlongView<T> pv,
          const int* __restrict__ done) {
  if (done && *done) return;
  const int64_t nx = op.nx, ny = op.ny;
  const int64_t ix = blockIdx.x * (int64_t)kMgThreads + threadIdx.x;
  if (ix >= nx) return;
  const bool xs = in_strip(ix, nx, npx);
  const int64_t ixm = ix == 0 ? nx - 1 : ix - 1, ixp = ix + 1 == nx ? 0 : ix + 1;
#pragma unroll
  for (int r = 0; r < kMgRows; ++r) {
    const int64_t iy = blockIdx.y * (int64_t)kMgRows + r;
    if (iy >= ny) break;
    const int64_t n = ix + nx * iy;
    const bool ys = ys_in(ysr, (int)iy);
    const int64_t iym = iy == 0 ? ny - 1 : iy - 1, iyp = iy + 1 == ny ? 0 : iy + 1;
    cplx<T> W, E, S, Nn, C;
    row_coefs<T, TE>(op, ix, iy, ixp, iyp, W, E, S, Nn, C);
    cplx<T> u0(T(0), T(0)), res = f[n];
    if (!ZERO) {
      u0 = iterate_at<T, PROLONG>(u, pv, nx, ny, ix, iy);
      res -= C * u0;
      res -= W * iterate_at<T, PROLONG>(u, pv, nx, ny, ixm, iy);
      res -= E * iterate_at<T, PROLONG>(u, pv, nx, ny, ixp, iy);
      res -= S * iterate_at<T, PROLONG>(u, pv, nx, ny, ix, iym);
      res -= Nn * iterate_at<T, PROLONG>(u, pv, nx, ny, ix, iyp);
    }
    if (xs) rxs[strip_line(ix, nx, npx) * ny + iy] = res;
    if (ys) rys[ys_line(ysr, (int)iy) * nx + ix] = res;
    out[n] = (xs || ys) ? u0 : u0 + wj * cdiv(res, C);
  }
}

// ---- shared-memory tiled variants for the bandwidth-bound fine levels ------------------------------------
// 32-bit index arithmetic throughout (a level never exceeds 2^31 points); the 64-bit version was issue bound.
// k_smooth2_tile: the (optionally coarse-corrected) iterate of a TX x TY tile plus a one-point halo is staged in
// shared memory once (the untiled kernel evaluates the interpolation 5x per point), then residual / Jacobi /
// strip capture run from the tile.  HBM traffic: u 8 + f 8 + mass 8 + out 8 B per point (fp32).
constexpr int kTX = 64, kTY = 16, kTileThreads = 256;

template <typename T, bool PROLONG>
__device__ __forceinline__ cplx<T> iterate32(const cplx<T>* __restrict__ u, const ProlongView<T>& pv, int nx, int ny, int ix, int iy) {
  cplx<T> v = u[ix + nx * iy];
  if (PROLONG) {
    const int nxc = (int)pv.nxc, nyc = (int)pv.nyc;
    const int I0 = ix >> 1, J0 = iy >> 1;
    const bool ox = ix & 1, oy = iy & 1;
    const int I1 = ox ? (I0 + 1 == nxc ? 0 : I0 + 1) : I0;
    const int J1 = oy ? (J0 + 1 == nyc ? 0 : J0 + 1) : J0;
    const cplx<T>* r0 = pv.uc + nxc * J0;
    cplx<T> lo = r0[I0];
    if (ox) lo = pv.pwx[ix] * lo + pv.pwx[nx + ix] * r0[I1];
    if (oy) {
      const cplx<T>* r1 = pv.uc + nxc * J1;
      cplx<T> hi = r1[I0];
      if (ox) hi = pv.pwx[ix] * hi + pv.pwx[nx + ix] * r1[I1];
      lo = pv.pwy[iy] * lo + pv.pwy[ny + iy] * hi;
    }
    v += lo;
  }
  return v;
}

template <typename T, bool TE, bool PROLONG>
__global__ void __launch_bounds__(kTileThreads)
k_smooth2_tile(OpView<T> op, const cplx<T>* __restrict__ u, const cplx<T>* __restrict__ f, cplx<T>* __restrict__ out,
               cplx<T>* __restrict__ rxs, cplx<T>* __restrict__ rys, int npx, YS ysr, T wj, ProlongView<T> pv,
               const int* __restrict__ done) {
  if (done && *done) return;
  __shared__ cplx<T> tile[(kTY + 2) * (kTX + 2)];
  const int nx = (int)op.nx, ny = (int)op.ny;
  const int x0 = blockIdx.x * kTX, y0 = blockIdx.y * kTY;
  for (int idx = threadIdx.x; idx < (kTX + 2) * (kTY + 2); idx += kTileThreads) {  // 1-D: every pass is full width
    const int lx = idx % (kTX + 2), ly = idx / (kTX + 2);
    int gx = x0 + lx - 1, gy = y0 + ly - 1;
    if (gx > nx || gy > ny) continue;  // beyond the wrap halo of a partial tile: never read
    gx = gx < 0 ? nx - 1 : (gx >= nx ? gx - nx : gx);
    gy = gy < 0 ? ny - 1 : (gy >= ny ? gy - ny : gy);
    tile[idx] = iterate32<T, PROLONG>(u, pv, nx, ny, gx, gy);
  }
  __syncthreads();
  const int lx = threadIdx.x % kTX;
  const int ix = x0 + lx;
  if (ix >= nx) return;
  const bool xs = in_strip(ix, nx, npx);
  const int ixp = ix + 1 == nx ? 0 : ix + 1;
  const cplx<T> cW = op.cxm[ix], cE = op.cxp[ix];
#pragma unroll
  for (int k = 0; k < kTY / (kTileThreads / kTX); ++k) {
    const int ly = threadIdx.x / kTX + k * (kTileThreads / kTX);
    const int iy = y0 + ly;
    if (iy >= ny) break;
    const int n = ix + nx * iy;
    const bool ys = ys_in(ysr, (int)iy);
    cplx<T> W = cW, E = cE, S = op.cym[iy], Nn = op.cyp[iy], m;
    if (TE) {
      const int iyp = iy + 1 == ny ? 0 : iy + 1;
      W = W * op.gx[n]; E = E * op.gx[ixp + nx * iy]; S = S * op.gy[n]; Nn = Nn * op.gy[ix + nx * iyp];
      m = op.mass_const;
    } else m = op.mass[n];
    const cplx<T> C = m - W - E - S - Nn;
    const cplx<T>* t = tile + (ly + 1) * (kTX + 2) + (lx + 1);
    const cplx<T> u0 = t[0];
    cplx<T> res = f[n];
    res -= C * u0; res -= W * t[-1]; res -= E * t[1]; res -= S * t[-(kTX + 2)]; res -= Nn * t[kTX + 2];
    if (xs) rxs[(int)strip_line(ix, nx, npx) * ny + iy] = res;
    if (ys) rys[(int)ys_line(ysr, (int)iy) * nx + ix] = res;
    out[n] = (xs || ys) ? u0 : u0 + wj * cdiv(res, C);
  }
}

// ---- k_smooth3: column-marching smoother for the bandwidth-bound levels --------------------------------------------
// One thread per fine column, kS3R rows per CTA (like the fp64 stencil).  Everything that depends on the column only
// (x-parity, coarse column, x-interpolation weights, x coefficients) is computed once per thread; the row parity is static
// because a CTA starts on an even row, so the coarse-grid correction costs one x-interpolation per coarse row and one
// y-interpolation per odd fine row instead of a 2-D index decode + 4 loads per staged point (the 64x16 tile version spent
// ~130 instructions per point and was issue bound at 3.2 TB/s).  The y-neighbours of the corrected iterate stay in
// registers, the x-neighbours go through one shared-memory row per fine row, all global loads are issued before the
// single barrier.  HBM traffic: u 8 + f 8 + mass 8 + out 8 B per point (+ 2 B coarse).
constexpr int kS3T = 128;

// a / b for the Jacobi update: fp32 uses the hardware reciprocal approximation (2 ulp; it only damps a smoothing update)
__device__ __forceinline__ cplx<float> cdiv_fast(cplx<float> a, cplx<float> b) {
  const float d = __fdividef(1.0f, b.x * b.x + b.y * b.y);
  return cplx<float>((a.x * b.x + a.y * b.y) * d, (a.y * b.x - a.x * b.y) * d);
}
__device__ __forceinline__ cplx<double> cdiv_fast(cplx<double> a, cplx<double> b) { return cdiv(a, b); }

// interior CTA of the TM smoother: no periodic wrap, no partial rows / columns, no PML strip in its rows or columns.
// Straight-line code on constant row strides (the general path below spends > 2/3 of its instructions on index
// arithmetic, bounds and strip predicates).
template <typename T, bool PROLONG, int kS3R>
__device__ __forceinline__ void smooth3_interior(const OpView<T>& op, const cplx<T>* __restrict__ u, const cplx<T>* __restrict__ f,
                                                 cplx<T>* __restrict__ out, T wj, const ProlongView<T>& pv,
                                                 cplx<T> (*sv)[kS3T + 2], int x0, int y0) {
  const int nx = (int)op.nx, ny = (int)op.ny;
  const int t = threadIdx.x, ix = x0 + t;
  const size_t n0 = (size_t)y0 * nx + ix;
  const cplx<T>* up = u + n0 - nx;          // row y0 - 1
  const cplx<T>* fp = f + n0;
  const cplx<T>* mp = op.mass + n0;
  cplx<T> v[kS3R + 2], fr[kS3R], mr[kS3R];
#pragma unroll
  for (int q = 0; q < kS3R + 2; ++q) v[q] = up[(size_t)q * nx];
#pragma unroll
  for (int r = 0; r < kS3R; ++r) { fr[r] = fp[(size_t)r * nx]; mr[r] = mp[(size_t)r * nx]; }
  if (PROLONG) {
    const int nxc = (int)pv.nxc;
    const int ox = ix & 1;
    const cplx<T>* cp = pv.uc + (size_t)((y0 >> 1) - 1) * nxc + (ix >> 1);   // coarse row J0 - 1
    // branch-free x-interpolation: even columns use weights (1, 0) and read the same coarse point twice
    cplx<T> wxl(T(1), T(0)), wxr(T(0), T(0));
    if (ox) { wxl = pv.pwx[ix]; wxr = pv.pwx[nx + ix]; }
    cplx<T> cx[kS3R / 2 + 2];
#pragma unroll
    for (int j = 0; j < kS3R / 2 + 2; ++j) {
      const cplx<T> c0 = cp[(size_t)j * nxc], c1 = cp[(size_t)j * nxc + ox];
      cx[j] = ox ? wxl * c0 + wxr * c1 : c0;
    }
    const cplx<T>* wy = pv.pwy + (y0 - 1);
#pragma unroll
    for (int q = 0; q < kS3R + 2; ++q) {      // slot q <-> row y0 - 1 + q: odd rows are the even slots
      if ((q & 1) == 0) v[q] += wy[q] * cx[q >> 1] + wy[ny + q] * cx[(q >> 1) + 1];
      else v[q] += cx[(q + 1) >> 1];
    }
  }
#pragma unroll
  for (int r = 0; r < kS3R; ++r) sv[r][t + 1] = v[r + 1];
  if (t < 2 * kS3R) {
    const int side = t / kS3R, r = t - side * kS3R;
    sv[r][side ? kS3T + 1 : 0] = iterate32<T, PROLONG>(u, pv, nx, ny, side ? x0 + kS3T : x0 - 1, y0 + r);
  }
  __syncthreads();
  const cplx<T> W = op.cxm[ix], E = op.cxp[ix];
  const cplx<T> mWE = -W - E;
  cplx<T>* op_ = out + n0;
#pragma unroll
  for (int r = 0; r < kS3R; ++r) {
    const cplx<T> S = op.cym[y0 + r], Nn = op.cyp[y0 + r];
    const cplx<T> C = mr[r] - W - E - S - Nn;
    const cplx<T> u0 = v[r + 1];
    cplx<T> res = fr[r];
    res -= C * u0; res -= W * sv[r][t]; res -= E * sv[r][t + 2]; res -= S * v[r]; res -= Nn * v[r + 2];
    op_[(size_t)r * nx] = u0 + wj * cdiv_fast(res, C);
  }
  (void)mWE;
}


template <typename T, bool TE, bool PROLONG, int kS3R>
__global__ void __launch_bounds__(kS3T, kS3R == 8 ? 4 : 6)
k_smooth3(OpView<T> op, const cplx<T>* __restrict__ u, const cplx<T>* __restrict__ f, cplx<T>* __restrict__ out,
          cplx<T>* __restrict__ rxs, cplx<T>* __restrict__ rys, int npx, YS ysr, T wj, ProlongView<T> pv,
          const int* __restrict__ done) {
  if (done && *done) return;
  __shared__ cplx<T> sv[kS3R][kS3T + 2];
  const int nx = (int)op.nx, ny = (int)op.ny;
  const int x0 = blockIdx.x * kS3T, y0 = blockIdx.y * kS3R;
  if (!TE) {
    const bool interior = x0 >= 1 && x0 + kS3T + 1 <= nx && y0 >= 2 && y0 + kS3R + 2 <= ny &&
                          x0 >= npx && x0 + kS3T <= nx - npx &&
                          (y0 + kS3R <= ysr.a0 || y0 >= ysr.a1) && (y0 + kS3R <= ysr.b0 || y0 >= ysr.b1);
    if (interior) { smooth3_interior<T, PROLONG, kS3R>(op, u, f, out, wj, pv, sv, x0, y0); return; }
  }
  const int t = threadIdx.x, ix = x0 + t;
  const int ncols = min(kS3T, nx - x0);
  const bool col = t < ncols;
  cplx<T> v[kS3R + 2];   // own column: slot s <-> fine row y0 - 1 + s (periodic)
  cplx<T> fr[kS3R], mr[kS3R];
  if (col) {
    // plain loads first (f, mass and the stored iterate of rows y0 .. y0 + R)
#pragma unroll
    for (int r = 0; r <= kS3R; ++r) {
      const int iy = y0 + r;
      if (iy < ny) {
        const int n = ix + nx * iy;
        v[r + 1] = u[n];
        if (r < kS3R) { fr[r] = f[n]; mr[r] = TE ? op.mass_const : op.mass[n]; }
      }
    }
    if (y0 > 0) v[0] = u[ix + nx * (y0 - 1)];
    if (PROLONG) {
      const int nxc = (int)pv.nxc, nyc = (int)pv.nyc;
      const int I0 = ix >> 1;
      const bool ox = ix & 1;
      const int I1 = ox ? (I0 + 1 == nxc ? 0 : I0 + 1) : I0;
      cplx<T> wxl(T(1), T(0)), wxr(T(0), T(0));
      if (ox) { wxl = pv.pwx[ix]; wxr = pv.pwx[nx + ix]; }
      const int J0 = y0 >> 1;
      // x-interpolated coarse values of coarse rows J0 - 1 (slot 0) and J0 .. J0 + R/2 (slots 1..)
      cplx<T> cx[kS3R / 2 + 2];
#pragma unroll
      for (int j = 0; j < kS3R / 2 + 2; ++j) {
        int J = J0 - 1 + j;
        if (J < 0) J = 0;          // unused (the row below row 0 takes the dynamic path)
        J %= nyc;
        const cplx<T>* rw = pv.uc + nxc * J;
        cplx<T> c = rw[I0];
        if (ox) c = wxl * c + wxr * rw[I1];
        cx[j] = c;
      }
#pragma unroll
      for (int r = 0; r <= kS3R; ++r) {      // row y0 + r has the parity of r
        const int iy = y0 + r;
        if (iy < ny) {
          if (r & 1) v[r + 1] += pv.pwy[iy] * cx[(r >> 1) + 1] + pv.pwy[ny + iy] * cx[(r >> 1) + 2];
          else v[r + 1] += cx[(r >> 1) + 1];
        }
      }
      if (y0 > 0) v[0] += pv.pwy[y0 - 1] * cx[0] + pv.pwy[ny + y0 - 1] * cx[1];   // odd row 2 (J0 - 1) + 1
    }
    // periodic neighbours of the first / last row: any parity, generic path
    if (y0 == 0) v[0] = iterate32<T, PROLONG>(u, pv, nx, ny, ix, ny - 1);
#pragma unroll
    for (int r = 1; r <= kS3R; ++r) if (y0 + r == ny) v[r + 1] = iterate32<T, PROLONG>(u, pv, nx, ny, ix, 0);
#pragma unroll
    for (int r = 0; r < kS3R; ++r) sv[r][t + 1] = v[r + 1];
  }
  if (t < 2 * kS3R) {   // the two columns next to the CTA's columns
    const int side = t / kS3R, r = t - side * kS3R, iy = y0 + r;
    if (iy < ny) {
      const int gx = side ? (x0 + ncols == nx ? 0 : x0 + ncols) : (x0 == 0 ? nx - 1 : x0 - 1);
      sv[r][side ? ncols + 1 : 0] = iterate32<T, PROLONG>(u, pv, nx, ny, gx, iy);
    }
  }
  __syncthreads();
  if (!col) return;
  const bool xs = in_strip(ix, nx, npx);
  const int ixp = ix + 1 == nx ? 0 : ix + 1;
  const cplx<T> cW = op.cxm[ix], cE = op.cxp[ix];
#pragma unroll
  for (int r = 0; r < kS3R; ++r) {
    const int iy = y0 + r;
    if (iy >= ny) break;
    const int n = ix + nx * iy;
    const bool ys = ys_in(ysr, iy);
    cplx<T> W = cW, E = cE, S = op.cym[iy], Nn = op.cyp[iy];
    if (TE) {
      const int iyp = iy + 1 == ny ? 0 : iy + 1;
      W = W * op.gx[n]; E = E * op.gx[ixp + nx * iy]; S = S * op.gy[n]; Nn = Nn * op.gy[ix + nx * iyp];
    }
    const cplx<T> C = mr[r] - W - E - S - Nn;
    const cplx<T> u0 = v[r + 1];
    cplx<T> res = fr[r];
    res -= C * u0; res -= W * sv[r][t]; res -= E * sv[r][t + 2]; res -= S * v[r]; res -= Nn * v[r + 2];
    if (xs) rxs[(int)strip_line(ix, nx, npx) * ny + iy] = res;
    if (ys) rys[(int)ys_line(ysr, iy) * nx + ix] = res;
    out[n] = (xs || ys) ? u0 : u0 + wj * cdiv(res, C);
  }
}

// k_restrict_tile: residual of a (2 CX + 1) x (2 CY + 1) fine patch computed once into shared memory, then the
// CX x CY coarse points of the tile take their 3 x 3 weighted sums.  HBM traffic: u 8 + f 8 + mass 8 B per fine
// point + 2 B write (fp32) instead of 2.25x recomputation with stride-2 access.
constexpr int kCX = 32, kCY = 8;
constexpr int kRW = 2 * kCX + 1, kRH = 2 * kCY + 1;   // residual patch
constexpr int kUW = kRW + 2, kUH = kRH + 2;           // iterate patch

__device__ __forceinline__ int wrap32(int i, int n) {  // cheap for |i - [0,n)| < n, correct for tiny levels too
  while (i < 0) i += n;
  while (i >= n) i -= n;
  return i;
}

template <typename T, bool TE>
__global__ void __launch_bounds__(kTileThreads)
k_restrict_tile(OpView<T> op, const cplx<T>* __restrict__ u, const cplx<T>* __restrict__ f, int64_t nxc64, int64_t nyc64,
                const cplx<T>* __restrict__ rx, const cplx<T>* __restrict__ ry, cplx<T>* __restrict__ fc,
                const int* __restrict__ done) {
  if (done && *done) return;
  __shared__ cplx<T> su[kUW * kUH];
  __shared__ cplx<T> sr[kRW * kRH];
  __shared__ int sgx[kUW], sgy[kUH];
  __shared__ cplx<T> scxm[kUW], scxp[kUW], scym[kUH], scyp[kUH];  // 1-D coefficient slices of the patch (L1 relief)
  const int nx = (int)op.nx, ny = (int)op.ny, nxc = (int)nxc64, nyc = (int)nyc64;
  const int I0 = blockIdx.x * kCX, J0 = blockIdx.y * kCY;
  // fine index of iterate-patch slot (lx,ly): 2 I0 - 2 + lx, wrapped (plain modulo: slots stay periodic neighbours)
  const int fx0 = 2 * I0 - 2, fy0 = 2 * J0 - 2;
  if (threadIdx.x < kUW) {
    const int gx = wrap32(fx0 + threadIdx.x, nx);
    sgx[threadIdx.x] = gx; scxm[threadIdx.x] = op.cxm[gx]; scxp[threadIdx.x] = op.cxp[gx];
  } else if (threadIdx.x >= 128 && threadIdx.x < 128 + kUH) {
    const int gy = wrap32(fy0 + (threadIdx.x - 128), ny);
    sgy[threadIdx.x - 128] = gy; scym[threadIdx.x - 128] = op.cym[gy]; scyp[threadIdx.x - 128] = op.cyp[gy];
  }
  __syncthreads();
  {
    constexpr int NU = (kUW * kUH + kTileThreads - 1) / kTileThreads;
    cplx<T> reg[NU];
#pragma unroll
    for (int k = 0; k < NU; ++k) {   // all loads first (memory-level parallelism), then the shared-memory stores
      const int idx = min(threadIdx.x + k * kTileThreads, kUW * kUH - 1);
      const int lx = idx % kUW, ly = idx / kUW;
      reg[k] = u[sgx[lx] + nx * sgy[ly]];
    }
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      const int idx = threadIdx.x + k * kTileThreads;
      if (idx < kUW * kUH) su[idx] = reg[k];
    }
  }
  __syncthreads();
  {
    constexpr int NR = (kRW * kRH + kTileThreads - 1) / kTileThreads;
    cplx<T> fv[NR], mv[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int idx = min(threadIdx.x + k * kTileThreads, kRW * kRH - 1);
      const int lx = idx % kRW, ly = idx / kRW;
      const int n = sgx[lx + 1] + nx * sgy[ly + 1];
      fv[k] = f[n];
      mv[k] = TE ? op.mass_const : op.mass[n];
    }
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int idx = threadIdx.x + k * kTileThreads;
      if (idx >= kRW * kRH) break;
      const int lx = idx % kRW, ly = idx / kRW;
      cplx<T> W = scxm[lx + 1], E = scxp[lx + 1], S = scym[ly + 1], Nn = scyp[ly + 1];
      if (TE) {
        const int ix = sgx[lx + 1], iy = sgy[ly + 1];
        const int n = ix + nx * iy;
        const int ixp = ix + 1 == nx ? 0 : ix + 1, iyp = iy + 1 == ny ? 0 : iy + 1;
        W = W * op.gx[n]; E = E * op.gx[ixp + nx * iy]; S = S * op.gy[n]; Nn = Nn * op.gy[ix + nx * iyp];
      }
      const cplx<T> C = mv[k] - W - E - S - Nn;
      const cplx<T>* t = su + (ly + 1) * kUW + (lx + 1);
      cplx<T> res = fv[k];
      res -= C * t[0]; res -= W * t[-1]; res -= E * t[1]; res -= S * t[-kUW]; res -= Nn * t[kUW];
      sr[idx] = res;
    }
  }
  __syncthreads();
  const int ci = threadIdx.x % kCX, cj = threadIdx.x / kCX;
  const int I = I0 + ci, J = J0 + cj;
  if (I >= nxc || J >= nyc) return;
  cplx<T> acc(T(0), T(0));
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const cplx<T> wy = ry[3 * J + b];
#pragma unroll
    for (int a = 0; a < 3; ++a) acc += (rx[3 * I + a] * wy) * sr[(2 * cj + b) * kRW + (2 * ci + a)];
  }
  fc[I + nxc * J] = acc;
}

// ---- register-marching variants (fewest instructions per point) ------------------------------------------------
// One thread per fine column marches down the rows keeping the three y-neighbours of the (corrected) iterate in
// registers; x-neighbours come from warp shuffles, so lanes 0 and 31 of each warp are halo lanes (30 useful columns
// per warp).  Per point: one interpolation, one u / f / mass load, one store -- the shared-memory tile versions spent
// ~130 instructions per point (issue bound at 3.2 TB/s).
constexpr int kMW = 30, kMWarps = 4, kMRows = 32;

template <typename T> __device__ __forceinline__ cplx<T> shfl_up_c(cplx<T> v) {
  return cplx<T>(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}
template <typename T> __device__ __forceinline__ cplx<T> shfl_dn_c(cplx<T> v) {
  return cplx<T>(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}

template <typename T, bool TE, bool PROLONG>
__global__ void __launch_bounds__(32 * kMWarps)
k_smooth2_march(OpView<T> op, const cplx<T>* __restrict__ u, const cplx<T>* __restrict__ f, cplx<T>* __restrict__ out,
                cplx<T>* __restrict__ rxs, cplx<T>* __restrict__ rys, int npx, YS ysr, T wj, ProlongView<T> pv,
                const int* __restrict__ done) {
  if (done && *done) return;
  const int nx = (int)op.nx, ny = (int)op.ny;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col0 = (blockIdx.x * kMWarps + warp) * kMW;
  if (col0 >= nx) return;                      // whole warp idle (no block-level barriers in this kernel)
  const int ix = col0 + lane - 1;              // lanes 0 / 31: left / right halo column
  const bool valid = ix <= nx;                 // ix == -1 and ix == nx are periodic halos
  const bool useful = lane >= 1 && lane <= kMW && ix < nx;
  const int ixw = !valid ? 0 : (ix < 0 ? nx - 1 : (ix == nx ? 0 : ix));
  const int y0 = blockIdx.y * kMRows;
  const cplx<T> zero(T(0), T(0));
  cplx<T> cW = zero, cE = zero;
  bool xs = false;
  int ixp = 0;
  if (useful) { cW = op.cxm[ix]; cE = op.cxp[ix]; xs = in_strip(ix, nx, npx); ixp = ix + 1 == nx ? 0 : ix + 1; }
  const int iym0 = y0 == 0 ? ny - 1 : y0 - 1;
  cplx<T> vS = valid ? iterate32<T, PROLONG>(u, pv, nx, ny, ixw, iym0) : zero;
  cplx<T> vC = valid ? iterate32<T, PROLONG>(u, pv, nx, ny, ixw, y0) : zero;
  for (int r = 0; r < kMRows; ++r) {
    const int iy = y0 + r;
    if (iy >= ny) break;                       // uniform over the CTA
    const int iyp = iy + 1 == ny ? 0 : iy + 1;
    const cplx<T> vN = valid ? iterate32<T, PROLONG>(u, pv, nx, ny, ixw, iyp) : zero;
    const cplx<T> vW = shfl_up_c(vC), vE = shfl_dn_c(vC);
    if (useful) {
      const int n = ix + nx * iy;
      const bool ys = ys_in(ysr, (int)iy);
      cplx<T> W = cW, E = cE, S = op.cym[iy], Nn = op.cyp[iy], m;
      if (TE) {
        W = W * op.gx[n]; E = E * op.gx[ixp + nx * iy]; S = S * op.gy[n]; Nn = Nn * op.gy[ix + nx * iyp];
        m = op.mass_const;
      } else m = op.mass[n];
      const cplx<T> C = m - W - E - S - Nn;
      cplx<T> res = f[n];
      res -= C * vC; res -= W * vW; res -= E * vE; res -= S * vS; res -= Nn * vN;
      if (xs) rxs[(int)strip_line(ix, nx, npx) * ny + iy] = res;
      if (ys) rys[(int)ys_line(ysr, (int)iy) * nx + ix] = res;
      out[n] = (xs || ys) ? vC : vC + wj * cdiv(res, C);
    }
    vS = vC; vC = vN;
  }
}

// residual + restriction, marching: lanes 1..30 hold residual columns, even-ix lanes 2..28 own coarse points; the
// x-part of the (separable) restriction comes from shuffles of the residual, the y-part accumulates in registers over
// the three fine rows of a coarse row.  fine column of lane l: c0 + l - 2, c0 = 28 * warp index (even).
constexpr int kRWc = 28, kRCRows = 16;   // fine columns with a coarse owner per warp; coarse rows per CTA

template <typename T, bool TE>
__global__ void __launch_bounds__(32 * kMWarps)
k_restrict_march(OpView<T> op, const cplx<T>* __restrict__ u, const cplx<T>* __restrict__ f, int64_t nxc64, int64_t nyc64,
                 const cplx<T>* __restrict__ rx, const cplx<T>* __restrict__ ry, cplx<T>* __restrict__ fc,
                 const int* __restrict__ done) {
  if (done && *done) return;
  const int nx = (int)op.nx, ny = (int)op.ny, nxc = (int)nxc64, nyc = (int)nyc64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * kMWarps + warp) * kRWc;
  if (c0 >= nx) return;
  const int ix = c0 + lane - 2;                                   // in [-2, nx + 1]
  const bool valid = ix <= nx + 1;
  const int ixw = !valid ? 0 : wrap32(ix, nx);
  const bool has_res = lane >= 1 && lane <= 30 && valid;         // residual defined (both x-neighbours in the warp)
  const bool owner = lane >= 2 && lane <= 28 && (lane & 1) == 0 && ix < nx;  // even ix: coarse point I = ix / 2
  const int I = ix >> 1;
  const cplx<T> zero(T(0), T(0));
  cplx<T> cW = zero, cE = zero, w0 = zero, w1 = zero, w2 = zero;
  if (has_res) { cW = op.cxm[ixw]; cE = op.cxp[ixw]; }
  if (owner) { w0 = rx[3 * I]; w1 = rx[3 * I + 1]; w2 = rx[3 * I + 2]; }
  const int J0 = blockIdx.y * kRCRows;
  const int ixpw = ixw + 1 == nx ? 0 : ixw + 1;
  // fine rows 2 J0 - 1 ... 2 (J0 + CR - 1) + 1 ; iterate rows one further on each side
  int iy = wrap32(2 * J0 - 1, ny);
  cplx<T> vS = valid ? u[ixw + nx * wrap32(2 * J0 - 2, ny)] : zero;
  cplx<T> vC = valid ? u[ixw + nx * iy] : zero;
  cplx<T> acc = zero, acc_next = zero;
  for (int r = 0; r < 2 * kRCRows + 1; ++r) {
    const int J = J0 + (r >> 1);                 // r even: fine row 2J - 1 (bottom of J) ; r odd: fine row 2J
    if (J > nyc || (J == nyc && (r & 1))) break;  // uniform
    const int iyp = iy + 1 == ny ? 0 : iy + 1;
    const cplx<T> vN = valid ? u[ixw + nx * iyp] : zero;
    const cplx<T> vW = shfl_up_c(vC), vE = shfl_dn_c(vC);
    cplx<T> res = zero;
    if (has_res) {
      const int n = ixw + nx * iy;
      cplx<T> W = cW, E = cE, S = op.cym[iy], Nn = op.cyp[iy], m;
      if (TE) {
        W = W * op.gx[n]; E = E * op.gx[ixpw + nx * iy]; S = S * op.gy[n]; Nn = Nn * op.gy[ixw + nx * iyp];
        m = op.mass_const;
      } else m = op.mass[n];
      const cplx<T> C = m - W - E - S - Nn;
      res = f[n];
      res -= C * vC; res -= W * vW; res -= E * vE; res -= S * vS; res -= Nn * vN;
    }
    const cplx<T> rl = shfl_up_c(res), rr = shfl_dn_c(res);
    if (owner) {
      const cplx<T> rxrow = w0 * rl + w1 * res + w2 * rr;
      if ((r & 1) == 0) {                        // fine row 2J - 1: top of coarse J-1 (already added below), bottom of J
        if (J < nyc) acc_next = ry[3 * J] * rxrow;
        if (r > 0) {                             // completes coarse row J - 1
          acc += ry[3 * (J - 1) + 2] * rxrow;
          fc[I + nxc * (J - 1)] = acc;
        }
        acc = acc_next;
      } else {                                   // fine row 2J: centre of coarse J
        acc += ry[3 * J + 1] * rxrow;
      }
    }
    vS = vC; vC = vN; iy = iyp;
  }
}

constexpr int kLineMaxK = 10;   // segments are at most 640 points long

template <typename T>
__global__ void k_lines2(int64_t nx, int64_t ny, int npx, YS ys, LineSeg sy, LineSeg sx, const cplx<T>* __restrict__ mult_y,
                         const cplx<T>* __restrict__ mult_x, const cplx<T>* __restrict__ rxs, const cplx<T>* __restrict__ rys,
                         cplx<T>* __restrict__ out, T wl, int blk_off, unsigned long long* __restrict__ slots, const int* __restrict__ done) {
  if (done && *done) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nby = 2 * npx * sy.nseg;
  const int blk = (int)blockIdx.x + blk_off;
  const bool ymode = blk < nby;                       // y-line of a strip column, else x-line of a strip row
  const int b = ymode ? blk : blk - nby;
  const LineSeg sg = ymode ? sy : sx;
  const int c = b / sg.nseg, j = b % sg.nseg;
  const int lo = sg.lo(j), n = sg.hi(j) - lo, K = sg.K, SL = sg.SL;
  cplx<T>* d0 = reinterpret_cast<cplx<T>*>(smem_raw);
  cplx<T>* d1 = d0 + SL;
  const cplx<T>* alpha = (ymode ? mult_y : mult_x) + (size_t)b * (2 * K + 1) * SL;
  const cplx<T>* gamma = alpha + (size_t)K * SL;
  const cplx<T>* binv = gamma + (size_t)K * SL;
  const cplx<T>* rbuf = (ymode ? rxs : rys) + (size_t)c * sg.n + lo;
  const int i = threadIdx.x;              // one line point per thread
  const bool act = i < n;
  // every multiplier of this point is fetched up front (one load latency instead of one per PCR step)
  cplx<T> al[kLineMaxK], ga[kLineMaxK];
  cplx<T> bi(T(0), T(0));
  if (act) {
    d0[i] = rbuf[i];
#pragma unroll
    for (int k = 0; k < kLineMaxK; ++k) if (k < K) { al[k] = alpha[(size_t)k * SL + i]; ga[k] = gamma[(size_t)k * SL + i]; }
    bi = binv[i];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kLineMaxK; ++k) {
    if (k < K) {
      const int s = 1 << k;
      if (act) {
        cplx<T> v = d0[i];
        if (i >= s) v += al[k] * d0[i - s];
        if (i + s < n) v += ga[k] * d0[i + s];
        d1[i] = v;
      }
      __syncthreads();
      cplx<T>* t = d0; d0 = d1; d1 = t;
    }
  }
  const int gi = lo + i;                  // position on the line
  if (!act || gi < sg.core_lo(j) || gi >= sg.core_hi(j)) return;
  // corners (strip column AND strip row): the stretched operator is anisotropic in both directions there, so they take the
  // mean of their y-line and their x-line update (both from the same residual).  Measured (tools/gpu_mgdiag.py, 512^2):
  // with x-lines only in the corners the cycle's asymptotic factor is 0.955 and two sweeps per level diverge; with the
  // mean it is 0.45 and BiCGSTAB needs 92 instead of 167 iterations.
  // The two updates of a corner point come from different CTAs.  fp32: ONE launch -- the CTAs meet in a 64-bit slot per corner
  // point: whoever arrives first parks its update there with an atomic exchange, the second finds it, adds
  // u += w (y-update + x-update) in that fixed order (bit-reproducible whatever the arrival order, unlike two atomic adds) and
  // re-arms the slot.  fp64 (no 128-bit exchange): slots == NULL, y-lines and x-lines are two launches (blk_off).
  const bool corner = ymode ? ys_in(ys, gi) : in_strip(gi, nx, npx);
  const int fixed = ymode ? (int)strip_index(c, nx, npx) : ys_row(ys, c);
  const int64_t idx = ymode ? (int64_t)fixed + nx * (int64_t)gi : (int64_t)gi + nx * (int64_t)fixed;
  const cplx<T> upd = d0[i] * bi;
  if (!corner) { out[idx] += wl * upd; return; }
  const T w = T(0.5) * wl;
  if (sizeof(T) != 4 || slots == nullptr) { out[idx] += w * upd; return; }
  const int nys = (ys.a1 - ys.a0) + (ys.b1 - ys.b0);
  const int xl = ymode ? c : (int)strip_line(gi, nx, npx);     // strip column number 0 .. 2 npx - 1
  const int yl = ymode ? ys_line(ys, gi) : c;                  // strip row number 0 .. nys - 1
  unsigned long long* slot = slots + (size_t)xl * nys + yl;
  constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;    // (NaN, NaN) with an all-ones payload: never produced by arithmetic
  unsigned long long mine = ((unsigned long long)__float_as_uint((float)upd.y) << 32) | (unsigned long long)__float_as_uint((float)upd.x);
  if (mine == kEmpty) mine = 0x7FC000007FC00000ull;             // keep the sentinel unique (a diverged iterate is NaN anyway)
  const unsigned long long other = atomicExch(slot, mine);
  if (other == kEmpty) return;                                  // first to arrive: the partner finishes the update
  const cplx<T> o(T(__uint_as_float((unsigned)(other & 0xFFFFFFFFull))), T(__uint_as_float((unsigned)(other >> 32))));
  const cplx<T> yu = ymode ? upd : o, xu = ymode ? o : upd;
  out[idx] += w * (yu + xu);
  *slot = kEmpty;
}

// ---- residual + restriction (coarse-point-centric).  r_c(I,J) = sum RX[I][a] RY[J][b] r(xi[a], yi[b]) with
// RX = P^T V / (2 V_c): transpose of the operator-dependent interpolation, weighted by the stretched cell volumes
template <typename T, bool TE>
__global__ void k_resid_restrict(OpView<T> op, const cplx<T>* __restrict__ u, const cplx<T>* __restrict__ f,
                                 int64_t nxc, int64_t nyc, const cplx<T>* __restrict__ rx, const cplx<T>* __restrict__ ry,
                                 cplx<T>* __restrict__ fc, const int* __restrict__ done) {
  if (done && *done) return;
  const int64_t I = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t J = blockIdx.y;
  if (I >= nxc) return;
  int64_t xi[3], yi[3];
  tr_index(I, op.nx, xi); tr_index(J, op.ny, yi);
  cplx<T> acc(T(0), T(0));
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const cplx<T> wy = ry[3 * J + b];
    if (wy.x == T(0) && wy.y == T(0)) continue;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const cplx<T> wx = rx[3 * I + a];
      if (wx.x == T(0) && wx.y == T(0)) continue;
      acc += (wx * wy) * residual_at<T, TE>(op, u, f, xi[a], yi[b], nullptr);
    }
  }
  fc[I + nxc * J] = acc;
}

// ---- operator-dependent prolongation + correction: odd fine points interpolate with the complex weights
// wl/wr = conductance-weighted (linear in the STRETCHED coordinate), even points inject
template <typename T>
__global__ void k_prolong_add(int64_t nx, int64_t ny, int64_t nxc, int64_t nyc, const cplx<T>* __restrict__ pwx,
                              const cplx<T>* __restrict__ pwy, const cplx<T>* __restrict__ uc, cplx<T>* __restrict__ u,
                              const int* __restrict__ done) {
  if (done && *done) return;
  const int64_t ix = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t iy = blockIdx.y;
  if (ix >= nx) return;
  const int64_t I0 = ix >> 1, J0 = iy >> 1;
  const bool ox = ix & 1, oy = iy & 1;
  const int64_t I1 = ox ? (I0 + 1 == nxc ? 0 : I0 + 1) : I0;
  const int64_t J1 = oy ? (J0 + 1 == nyc ? 0 : J0 + 1) : J0;
  const cplx<T> one(T(1), T(0)), zero(T(0), T(0));
  const cplx<T> wxl = ox ? pwx[ix] : one, wxr = ox ? pwx[nx + ix] : zero;
  const cplx<T> wyl = oy ? pwy[iy] : one, wyr = oy ? pwy[ny + iy] : zero;
  cplx<T> lo = wxl * uc[I0 + nxc * J0];
  if (ox) lo += wxr * uc[I1 + nxc * J0];
  cplx<T> v = wyl * lo;
  if (oy) {
    cplx<T> hi = wxl * uc[I0 + nxc * J1];
    if (ox) hi += wxr * uc[I1 + nxc * J1];
    v += wyr * hi;
  }
  u[ix + nx * iy] += v;
}

}  // namespace

// ---- host: Galerkin-consistent 1-D hierarchy in the stretched coordinate -------------------------------------
// E[i] = stretched length (in cell units) of the edge between points i-1 and i, V[i] = stretched volume of point i.
//   f.b ordering (driven.jl:35):      E = s_backward,            V = s_forward
//   b.f ordering (modulation.jl:82):  E[i] = s_forward[i-1],     V = s_backward
// row i of the 1-D operator:  (scale/h^2) / V[i] * ( (u[i-1]-u[i])/E[i] + (u[i+1]-u[i])/E[i+1] ).
// Coarsening i = 2I: interpolation weights from the edge conductances (linear in the stretched coordinate),
// coarse edge = sum of its fine edges, coarse volume = P^T V, restriction = P^T V / (2 V_c).  Odd sizes leave one
// short coarse edge between the last and the first point (exact, no approximation).
using cdh = std::complex<double>;
struct Hier1D {
  std::vector<std::vector<cdh>> E, V, wl, wr, R;  // per level; R[l] restricts level l -> l+1 (3 per coarse point)
};
static void build_hier1d(const std::vector<cdh>& E0, const std::vector<cdh>& V0, int nlev, Hier1D& H) {
  H.E.assign(1, E0); H.V.assign(1, V0); H.wl.clear(); H.wr.clear(); H.R.clear();
  for (int l = 0; l < nlev; ++l) {
    const std::vector<cdh>& E = H.E[l]; const std::vector<cdh>& V = H.V[l];
    const int64_t n = (int64_t)E.size();
    std::vector<cdh> wl(n), wr(n);
    for (int64_t i = 0; i < n; ++i) {
      const cdh a = 1.0 / E[i], ap = 1.0 / E[(i + 1) % n];
      wl[i] = a / (a + ap); wr[i] = ap / (a + ap);
    }
    H.wl.push_back(wl); H.wr.push_back(wr);
    if (l == nlev - 1) break;
    const int64_t nc = (n + 1) / 2;
    std::vector<cdh> Ec(nc), Vc(nc), R(3 * nc);
    for (int64_t I = 0; I < nc; ++I) {
      const int64_t i = 2 * I;
      const bool has_l = !(I == 0 && (n % 2) == 1) && n > 1;
      const bool has_r = i + 1 <= n - 1;
      const int64_t il = (i - 1 + n) % n, ir = has_r ? i + 1 : 0;
      const cdh nl = has_l ? wr[il] * V[il] : cdh(0, 0), nc0 = V[i], nr = has_r ? wl[ir] * V[ir] : cdh(0, 0);
      Vc[I] = (nl + nc0 + nr) / 2.0;
      R[3 * I + 0] = nl / (2.0 * Vc[I]); R[3 * I + 1] = nc0 / (2.0 * Vc[I]); R[3 * I + 2] = nr / (2.0 * Vc[I]);
      Ec[I] = (E[i] + (has_l ? E[il] : cdh(0, 0))) / 2.0;
    }
    H.R.push_back(R); H.E.push_back(Ec); H.V.push_back(Vc);
  }
}
static void hier_coefs(const Hier1D& H, int l, double scale_over_h2, std::vector<cdh>& cm, std::vector<cdh>& cp) {
  const std::vector<cdh>& E = H.E[l]; const std::vector<cdh>& V = H.V[l];
  const int64_t n = (int64_t)E.size();
  cm.resize(n); cp.resize(n);
  for (int64_t i = 0; i < n; ++i) { cm[i] = scale_over_h2 / (V[i] * E[i]); cp[i] = scale_over_h2 / (V[i] * E[(i + 1) % n]); }
}

// level sizes of the hierarchy on a (global) grid; force_levels > 0 fixes the depth (slab mode)
std::vector<std::pair<int64_t, int64_t>> mg_level_sizes(const fdfd_grid_t& g, double omega, const MGParams& prm, int force_levels) {
  const double eps0 = kEps0 * g.L0, mu0 = kMu0 * g.L0;
  std::vector<std::pair<int64_t, int64_t>> sizes;
  int64_t nx = g.Nx, ny = g.Ny;
  sizes.push_back({nx, ny});
  // stop coarsening once the coarsest level is mass dominated even in vacuum (k0 h >= kh_stop): there damped Jacobi
  // alone converges (|kappa| >> 4) and deeper levels would only add latency-bound launches
  const double k0 = omega * std::sqrt(eps0 * mu0);
  while ((int)sizes.size() < prm.max_levels) {
    if (force_levels > 0 && (int)sizes.size() >= force_levels) break;
    const double hl = std::min(grid_dx(g), grid_dy(g)) * (double)((int64_t)1 << (sizes.size() - 1));
    if (force_levels <= 0 && prm.kh_stop > 0 && k0 * hl >= prm.kh_stop) break;
    const int64_t cx = (nx + 1) / 2, cy = (ny + 1) / 2;
    if (cx < prm.min_n || cy < prm.min_n || cx < 2 || cy < 2) break;
    nx = cx; ny = cy; sizes.push_back({nx, ny});
  }
  return sizes;
}

// ==============================================================================================
template <typename T> int Multigrid<T>::setup(fdfd_ctx* ctx_, const FineOp& op, const MGParams& prm_, int first_level,
                                               const c128* eps_first) {
  ctx = ctx_; prm = prm_; te = op.pol == FDFD_TE; first = first_level;
  const fdfd_grid_t& g = op.g;
  const double eps0 = kEps0 * g.L0, mu0 = kMu0 * g.L0;
  const double scale = te ? 1.0 : 1.0 / mu0;
  // store M scaled to O(1) coefficients (TE couplings are ~1e21 in SI-normalised units: |C|^2 overflows fp32)
  rhs_scale = (te ? eps0 : 1.0) / std::abs(op.hc.cxm[g.Nx / 2]);
  lv.clear();
  // level sizes (slab: the depth was decided on the global grid; local rows = owned rows + 2 H_l halo rows, H_l = 2^(L-1-l),
  // so that local coarse row lc <-> local fine row 2 lc exactly like on a whole grid)
  const SlabInfo& sl = op.slab;
  const fdfd_grid_t& gg = sl.on ? sl.gg : g;   // global grid: PML profile, strips
  std::vector<std::pair<int64_t, int64_t>> sizes = mg_level_sizes(gg, op.omega, prm, sl.on ? sl.nlevels : 0);
  const int nlev = (int)sizes.size();
  auto Hl = [&](int l) -> int64_t { return sl.on ? (sl.H >> l) : 0; };
  std::vector<int64_t> nyg(nlev);              // global rows per level
  for (int l = 0; l < nlev; ++l) { nyg[l] = sizes[l].second; if (sl.on) sizes[l].second = (sl.nyl >> l) + 2 * Hl(l); }
  // slice a global per-row array of level l (per entries per row) to the slab's local rows (periodic wrap)
  auto ysl = [&](const std::vector<cdh>& v, int l, int per) -> std::vector<cdh> {
    if (!sl.on) return v;
    const int64_t nloc = sizes[l].second, n = nyg[l], off = (sl.y0 >> l) - Hl(l);
    std::vector<cdh> o((size_t)nloc * per);
    for (int64_t i = 0; i < nloc; ++i) {
      const int64_t gi = ((off + i) % n + n) % n;
      for (int k = 0; k < per; ++k) o[(size_t)i * per + k] = v[(size_t)gi * per + k];
    }
    return o;
  };
  lv.resize(sizes.size());
  // 1-D hierarchies from the reference s-factors (not inverted) at the PML frequency
  Hier1D HX, HY;
  {
    std::vector<cdh> sxf, sxb, syf, syb;
    host_sfactor(gg, 0, 1, op.omega_pml, sxf); host_sfactor(gg, 0, 0, op.omega_pml, sxb);
    host_sfactor(gg, 1, 1, op.omega_pml, syf); host_sfactor(gg, 1, 0, op.omega_pml, syb);
    auto EV = [&](const std::vector<cdh>& sf, const std::vector<cdh>& sb, std::vector<cdh>& E, std::vector<cdh>& V) {
      const int64_t n = (int64_t)sf.size();
      E.resize(n); V.resize(n);
      for (int64_t i = 0; i < n; ++i) {
        if (op.ordering == FDFD_ORDER_FB) { E[i] = sb[i]; V[i] = sf[i]; }
        else { E[i] = sf[(i + n - 1) % n]; V[i] = sb[i]; }
      }
    };
    std::vector<cdh> Ex, Vx, Ey, Vy;
    EV(sxf, sxb, Ex, Vx); EV(syf, syb, Ey, Vy);
    build_hier1d(Ex, Vx, (int)lv.size(), HX); build_hier1d(Ey, Vy, (int)lv.size(), HY);
  }
  size_t scratch_need = 0;
  for (size_t l = 0; l < lv.size(); ++l) {
    MGLevel<T>& L = lv[l];
    L.nx = sizes[l].first; L.ny = sizes[l].second; L.stride = (int64_t)1 << l;
    const int64_t N = L.nx * L.ny;
    if ((int)l < first_level) continue;   // partial hierarchy (levels >= first_level only): nothing resident above it
    // 1-D coefficients
    Coef1D hc;
    if (l == 0) hc = op.hc;
    else {
      const double hx = grid_dx(g) * (double)L.stride, hy = grid_dy(g) * (double)L.stride;
      hier_coefs(HX, (int)l, scale / (hx * hx), hc.cxm, hc.cxp);
      hier_coefs(HY, (int)l, scale / (hy * hy), hc.cym, hc.cyp);
      hc.cym = ysl(hc.cym, (int)l, 1); hc.cyp = ysl(hc.cyp, (int)l, 1);
    }
    L.hc = hc;
    {  // transfer weights: prolongation from level l+1 (fine-indexed wl|wr) and restriction to level l+1
      std::vector<cplx<T>> pw; pw.reserve(2 * L.nx + 2 * L.ny);
      const std::vector<cdh> wly = ysl(HY.wl[l], (int)l, 1), wry = ysl(HY.wr[l], (int)l, 1);
      for (const std::vector<cdh>* v : {(const std::vector<cdh>*)&HX.wl[l], (const std::vector<cdh>*)&HX.wr[l], &wly, &wry}) for (auto& z : *v) pw.push_back(cplx<T>(T(z.real()), T(z.imag())));
      CUDA_TRY(ctx, L.pw.alloc(pw.size()));
      CUDA_TRY(ctx, cudaMemcpyAsync(L.pw.p, pw.data(), pw.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice, ctx->stream));
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      if (l + 1 < lv.size()) {
        std::vector<cplx<T>> rw; std::vector<c128> rwd;
        const std::vector<cdh> Ry = ysl(HY.R[l], (int)l + 1, 3);
        for (const std::vector<cdh>* v : {(const std::vector<cdh>*)&HX.R[l], &Ry}) for (auto& z : *v) { rw.push_back(cplx<T>(T(z.real()), T(z.imag()))); rwd.push_back(c128(z.real(), z.imag())); }
        CUDA_TRY(ctx, L.rw.alloc(rw.size())); CUDA_TRY(ctx, L.rwd.alloc(rwd.size()));
        CUDA_TRY(ctx, cudaMemcpyAsync(L.rw.p, rw.data(), rw.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(L.rwd.p, rwd.data(), rwd.size() * sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      }
    }
    std::vector<cplx<T>> pack; pack.reserve(2 * L.nx + 2 * L.ny);
    for (auto* v : {&hc.cxm, &hc.cxp, &hc.cym, &hc.cyp}) for (auto& z : *v) pack.push_back(cplx<T>(T(z.real() * rhs_scale), T(z.imag() * rhs_scale)));
    CUDA_TRY(ctx, L.c1d.alloc(pack.size()));
    CUDA_TRY(ctx, cudaMemcpyAsync(L.c1d.p, pack.data(), pack.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    // eps of this level
    const c128* eps_l = nullptr;
    const int threads = 256;
    if ((int)l == first_level && first_level > 0) eps_l = eps_first;
    else if (l == 0) eps_l = op.eps.p;
    else {
      CUDA_TRY(ctx, L.eps.alloc(N));
      const c128* eps_f = ((int)l - 1 == first_level && first_level > 0) ? eps_first : (l == 1 ? op.eps.p : lv[l - 1].eps.p);
      const int blocks = (int)std::min<int64_t>((N + threads - 1) / threads, (int64_t)ctx->num_sms * 16);
      k_restrict_eps<<<blocks, threads, 0, ctx->stream>>>(lv[l - 1].nx, lv[l - 1].ny, L.nx, L.ny, lv[l - 1].rwd.p,
                                                        lv[l - 1].rwd.p + 3 * L.nx, eps_f, L.eps.p);
      KLAUNCH(ctx);
      eps_l = L.eps.p;
    }
    {
      const int blocks = (int)std::min<int64_t>((N + threads - 1) / threads, (int64_t)ctx->num_sms * 16);
      if (te) {
        CUDA_TRY(ctx, L.gx.alloc(N)); CUDA_TRY(ctx, L.gy.alloc(N));
        k_level_setup<T, true><<<blocks, threads, 0, ctx->stream>>>(L.nx, L.ny, eps_l, 0.0, prm.beta, eps0, 0.0, nullptr, L.gx.p, L.gy.p);
        const double m = op.omega * op.omega * mu0 * rhs_scale;
        L.mass_const = cplx<T>(T(m), T(-prm.beta * m));
      } else {
        CUDA_TRY(ctx, L.mass.alloc(N));
        const double w2 = op.omega * op.omega * eps0 * rhs_scale;  // ~ k0^2 dx^2
        k_level_setup<T, false><<<blocks, threads, 0, ctx->stream>>>(L.nx, L.ny, eps_l, w2, prm.beta, eps0,
                                                                   prm.shift_growth * w2 * (double)(L.stride * L.stride),
                                                                   L.mass.p, nullptr, nullptr);
      }
      KLAUNCH(ctx);
      CUDA_TRY(ctx, cudaGetLastError());
    }
    CUDA_TRY(ctx, L.u.alloc(N)); CUDA_TRY(ctx, L.f.alloc(N)); CUDA_TRY(ctx, L.tmp.alloc(N));
    // PML strips of this level
    auto strip = [&](int64_t npml, int64_t n) -> int {
      if (npml == 0) return 0;
      int64_t w = (npml + L.stride - 1) / L.stride + prm.pad;
      return (int)std::min<int64_t>(w, n / 2);
    };
    L.npx = strip(gg.Npml_x, L.nx);
    if (!sl.on) { L.npy = strip(gg.Npml_y, L.ny); L.ys = YS{0, L.npy, (int)L.ny - L.npy, (int)L.ny}; }
    else {
      // every LOCAL row (halo rows included) whose global row lies in the y-strip [0,npy) U [NyG-npy,NyG) gets the x-line
      // relaxation, so a halo row is smoothed exactly as its owner smooths it (x-lines are never cut).  With the periodic
      // closure the strip rows form at most two runs of local rows (around the low and the high end of the slab).
      const int npyg = strip(gg.Npml_y, nyg[l]);
      const int64_t nloc = L.ny, n = nyg[l], off = (sl.y0 >> l) - Hl((int)l);
      int runs[3][2]; int nr = 0; bool in = false;
      for (int64_t i = 0; i <= nloc; ++i) {
        const int64_t gi = ((off + i) % n + n) % n;
        const bool s = i < nloc && npyg > 0 && (gi < npyg || gi >= n - npyg);
        if (s && !in) { if (nr < 3) runs[nr][0] = (int)i; in = true; }
        if (!s && in) { if (nr < 3) runs[nr][1] = (int)i; ++nr; in = false; }
      }
      ARG_CHECK(ctx, nr <= 2, "internal: the y-strip rows of a slab form more than two runs");
      L.npy = 0;
      L.ys = YS{nr > 0 ? runs[0][0] : 0, nr > 0 ? runs[0][1] : 0, nr > 1 ? runs[1][0] : 0, nr > 1 ? runs[1][1] : 0};
    }
    auto make_seg = [](int64_t n) {
      LineSeg g; g.n = (int)n;
      int Lc = 256, O = 128;
      if (const char* e = getenv("FDFD_MG_LINE_SEG")) { const int v = atoi(e); if (v <= 0) Lc = 1 << 30; else { Lc = v; O = v / 2; } }  // diagnostics
      if (n <= 640 || Lc + 2 * O > 640 || Lc >= n) { g.SL = (int)n; g.Lc = (int)n; g.O = 0; g.nseg = 1; }
      else { g.Lc = Lc; g.O = O; g.SL = Lc + 2 * O; g.nseg = (int)((n + Lc - 1) / Lc); }
      g.K = 0; while ((1 << g.K) < g.SL) ++g.K;
      return g;
    };
    L.sy = make_seg(L.ny); L.sx = make_seg(L.nx);
    if (L.npx > 0) {
      ARG_CHECK(ctx, L.sy.SL <= 1024 && L.sy.K <= 10 || L.sy.nseg > 1, "PML lines longer than 640 points need the segmented relaxation");
      CUDA_TRY(ctx, L.rxs.alloc((size_t)2 * L.npx * L.ny));
      CUDA_TRY(ctx, L.pcr_y.alloc((size_t)2 * L.npx * L.sy.nseg * (2 * L.sy.K + 1) * L.sy.SL));
      scratch_need = std::max(scratch_need, (size_t)2 * L.npx * L.sy.nseg * 6 * L.sy.SL);
    }
    if (L.npx > 0 && L.ys.count() > 0 && sizeof(T) == 4) {
      CUDA_TRY(ctx, L.corner_slots.alloc((size_t)2 * L.npx * L.ys.count()));
      CUDA_TRY(ctx, cudaMemsetAsync(L.corner_slots.p, 0xFF, (size_t)2 * L.npx * L.ys.count() * sizeof(unsigned long long), ctx->stream));
    }
    if (L.ys.count() > 0) {
      CUDA_TRY(ctx, L.rys.alloc((size_t)L.ys.count() * L.nx));
      CUDA_TRY(ctx, L.pcr_x.alloc((size_t)L.ys.count() * L.sx.nseg * (2 * L.sx.K + 1) * L.sx.SL));
      scratch_need = std::max(scratch_need, (size_t)L.ys.count() * L.sx.nseg * 6 * L.sx.SL);
    }
  }
  if (first_level == 0) CUDA_TRY(ctx, spare.alloc((size_t)g.Nx * g.Ny));
  CUDA_TRY(ctx, pcr_scratch.alloc(scratch_need));
  for (size_t l = 0; l < lv.size(); ++l) {
    MGLevel<T>& L = lv[l];
    if (L.npx > 0) {
      if (te) k_pcr_setup<T, true><<<2 * L.npx * L.sy.nseg, 256, 0, ctx->stream>>>(L.view(), 0, L.npx, L.ys, L.sy, pcr_scratch.p, L.pcr_y.p);
      else k_pcr_setup<T, false><<<2 * L.npx * L.sy.nseg, 256, 0, ctx->stream>>>(L.view(), 0, L.npx, L.ys, L.sy, pcr_scratch.p, L.pcr_y.p);
      KLAUNCH(ctx);
    }
    if (L.ys.count() > 0) {
      if (te) k_pcr_setup<T, true><<<L.ys.count() * L.sx.nseg, 256, 0, ctx->stream>>>(L.view(), 1, L.npy, L.ys, L.sx, pcr_scratch.p, L.pcr_x.p);
      else k_pcr_setup<T, false><<<L.ys.count() * L.sx.nseg, 256, 0, ctx->stream>>>(L.view(), 1, L.npy, L.ys, L.sx, pcr_scratch.p, L.pcr_x.p);
      KLAUNCH(ctx);
    }
    CUDA_TRY(ctx, cudaGetLastError());
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  pcr_scratch.release();
  return FDFD_OK;
}

// timing diagnostics only (results are wrong): FDFD_MG_SKIP bit 0 = no line relaxation, bit 1 = no tiled smoother,
// bit 2 = no restriction, bit 3 = no zero-guess smoother
static int mg_skip() { static const int v = []() { const char* e = getenv("FDFD_MG_SKIP"); return e ? atoi(e) : 0; }(); return v; }

template <typename T> int Multigrid<T>::smooth(int l, bool zero, bool prolong) {
  MGLevel<T>& L = lv[l];
  const OpView<T> op = L.view();
  dim3 grid((unsigned)((L.nx + kMgThreads - 1) / kMgThreads), (unsigned)((L.ny + kMgRows - 1) / kMgRows));
  const T wj = T(prm.wjac), wl = T(prm.wline);
  cplx<T>* out = zero ? L.u.p : L.tmp.p;
  ProlongView<T> pv{0, 0, nullptr, nullptr, nullptr};
  if (prolong) { MGLevel<T>& C = lv[l + 1]; pv = ProlongView<T>{C.nx, C.ny, L.pw.p, L.pw.p + 2 * L.nx, C.u.p}; }
  static const bool use_tile = []() { const char* e = getenv("FDFD_MG_KERNELS"); return !(e && std::string(e) == "march"); }();
  static const bool use_col = []() { const char* e = getenv("FDFD_MG_KERNELS"); return !e || std::string(e) == "col"; }();
  static const int s3r = []() { const char* e = getenv("FDFD_MG_S3R"); return e && atoi(e) == 4 ? 4 : 8; }();
  dim3 cgrid((unsigned)((L.nx + kS3T - 1) / kS3T), (unsigned)((L.ny + s3r - 1) / s3r));
  dim3 tgrid((unsigned)((L.nx + kTX - 1) / kTX), (unsigned)((L.ny + kTY - 1) / kTY));
  dim3 mgrid((unsigned)((L.nx + kMW * kMWarps - 1) / (kMW * kMWarps)), (unsigned)((L.ny + kMRows - 1) / kMRows));
#define SM2(TEV, ZV, PV) k_smooth2<T, TEV, ZV, PV><<<grid, kMgThreads, 0, ctx->stream>>>(op, L.u.p, L.f.p, out, L.rxs.p, L.rys.p, L.npx, L.ys, wj, pv, done)
#define SMT(TEV, PV) do { if (use_col && s3r == 8) k_smooth3<T, TEV, PV, 8><<<cgrid, kS3T, 0, ctx->stream>>>(op, L.u.p, L.f.p, out, L.rxs.p, L.rys.p, L.npx, L.ys, wj, pv, done); \
    else if (use_col) k_smooth3<T, TEV, PV, 4><<<cgrid, kS3T, 0, ctx->stream>>>(op, L.u.p, L.f.p, out, L.rxs.p, L.rys.p, L.npx, L.ys, wj, pv, done); \
    else if (use_tile) k_smooth2_tile<T, TEV, PV><<<tgrid, kTileThreads, 0, ctx->stream>>>(op, L.u.p, L.f.p, out, L.rxs.p, L.rys.p, L.npx, L.ys, wj, pv, done); \
    else k_smooth2_march<T, TEV, PV><<<mgrid, 32 * kMWarps, 0, ctx->stream>>>(op, L.u.p, L.f.p, out, L.rxs.p, L.rys.p, L.npx, L.ys, wj, pv, done); } while (0)
  if ((zero && (mg_skip() & 8)) || (!zero && (mg_skip() & 2))) {}
  else if (te) { if (zero) SM2(true, true, false); else if (prolong) SMT(true, true); else SMT(true, false); }
  else    { if (zero) SM2(false, true, false); else if (prolong) SMT(false, true); else SMT(false, false); }
#undef SM2
#undef SMT
  KLAUNCH(ctx);
  if (!zero) std::swap(L.u.p, L.tmp.p);
  const int nby = 2 * L.npx * L.sy.nseg, nbx = L.ys.count() * L.sx.nseg;
  if (!(mg_skip() & 1) && nby + nbx > 0) {
    if (L.corner_slots.p || nby == 0 || nbx == 0) {
      // one launch for all lines (the corner points' two updates meet in L.corner_slots, see k_lines2)
      const int slmax = std::max(nby > 0 ? L.sy.SL : 0, nbx > 0 ? L.sx.SL : 0);
      const int threads = std::max(32, ((slmax + 31) / 32) * 32);
      k_lines2<T><<<nby + nbx, threads, (size_t)2 * slmax * sizeof(cplx<T>), ctx->stream>>>(L.nx, L.ny, L.npx, L.ys, L.sy, L.sx, L.pcr_y.p, L.pcr_x.p,
                                                                                         L.rxs.p, L.rys.p, L.u.p, wl, 0, L.corner_slots.p, done);
      KLAUNCH(ctx);
    } else {
      // fp64 multigrid: y-lines first, then x-lines, so that a corner point is never updated by two CTAs of one launch
      const int ty = std::max(32, ((L.sy.SL + 31) / 32) * 32), tx = std::max(32, ((L.sx.SL + 31) / 32) * 32);
      k_lines2<T><<<nby, ty, (size_t)2 * L.sy.SL * sizeof(cplx<T>), ctx->stream>>>(L.nx, L.ny, L.npx, L.ys, L.sy, L.sx, L.pcr_y.p, L.pcr_x.p,
                                                                                L.rxs.p, L.rys.p, L.u.p, wl, 0, nullptr, done);
      KLAUNCH(ctx);
      k_lines2<T><<<nbx, tx, (size_t)2 * L.sx.SL * sizeof(cplx<T>), ctx->stream>>>(L.nx, L.ny, L.npx, L.ys, L.sy, L.sx, L.pcr_y.p, L.pcr_x.p,
                                                                                L.rxs.p, L.rys.p, L.u.p, wl, nby, nullptr, done);
      KLAUNCH(ctx);
    }
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return FDFD_OK;
}

template <typename T> int Multigrid<T>::restrict_residual(int l) {
  MGLevel<T>& L = lv[l];
  MGLevel<T>& C = lv[l + 1];
  {
    static const bool use_tile = []() { const char* e = getenv("FDFD_MG_KERNELS"); return !(e && std::string(e) == "march"); }();
    if (mg_skip() & 4) return FDFD_OK;
    if (use_tile) {
      dim3 grid((unsigned)((C.nx + kCX - 1) / kCX), (unsigned)((C.ny + kCY - 1) / kCY));
      if (te) k_restrict_tile<T, true><<<grid, kTileThreads, 0, ctx->stream>>>(L.view(), L.u.p, L.f.p, C.nx, C.ny, L.rw.p, L.rw.p + 3 * C.nx, C.f.p, done);
      else k_restrict_tile<T, false><<<grid, kTileThreads, 0, ctx->stream>>>(L.view(), L.u.p, L.f.p, C.nx, C.ny, L.rw.p, L.rw.p + 3 * C.nx, C.f.p, done);
    } else {
      dim3 grid((unsigned)((L.nx + kRWc * kMWarps - 1) / (kRWc * kMWarps)), (unsigned)((C.ny + kRCRows - 1) / kRCRows));
      if (te) k_restrict_march<T, true><<<grid, 32 * kMWarps, 0, ctx->stream>>>(L.view(), L.u.p, L.f.p, C.nx, C.ny, L.rw.p, L.rw.p + 3 * C.nx, C.f.p, done);
      else k_restrict_march<T, false><<<grid, 32 * kMWarps, 0, ctx->stream>>>(L.view(), L.u.p, L.f.p, C.nx, C.ny, L.rw.p, L.rw.p + 3 * C.nx, C.f.p, done);
    }
    KLAUNCH(ctx);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return FDFD_OK;
}

// kind: 0 = V, 1 = F, 2 = W (W recursion only while l < wdepth)
template <typename T> int Multigrid<T>::cycle(int l, bool zero, int kind) {
  MGLevel<T>& L = lv[l];
  if (l == (int)lv.size() - 1) {
    for (int s = 0; s < std::max(1, prm.coarse_sweeps); ++s) FDFD_TRY(smooth(l, zero && s == 0));
    return FDFD_OK;
  }
  for (int s = 0; s < std::max(1, prm.nu1); ++s) FDFD_TRY(smooth(l, zero && s == 0));
  MGLevel<T>& C = lv[l + 1];
  FDFD_TRY(restrict_residual(l));
  if (kind == 2 && l < wbase + prm.wdepth) {
    FDFD_TRY(cycle(l + 1, true, 2));
    FDFD_TRY(cycle(l + 1, false, 2));
  } else if (kind == 1) {
    FDFD_TRY(cycle(l + 1, true, 1));
    FDFD_TRY(cycle(l + 1, false, 0));
  } else {
    FDFD_TRY(cycle(l + 1, true, kind == 2 ? 0 : kind));
  }
  if (prm.nu2 < 1) {
    dim3 grid((unsigned)((L.nx + 127) / 128), (unsigned)L.ny);
    k_prolong_add<T><<<grid, 128, 0, ctx->stream>>>(L.nx, L.ny, C.nx, C.ny, L.pw.p, L.pw.p + 2 * L.nx, C.u.p, L.u.p, done);
    KLAUNCH(ctx);
  }
  for (int s = 0; s < prm.nu2; ++s) FDFD_TRY(smooth(l, false, s == 0));  // the first post-sweep applies the correction
  CUDA_TRY(ctx, cudaGetLastError());
  return FDFD_OK;
}

template <typename T> int Multigrid<T>::apply(const cplx<T>** out) {
  FDFD_TRY(cycle(0, true, prm.cycle));
  *out = lv[0].u.p;
  return FDFD_OK;
}

template struct Multigrid<float>;
template struct Multigrid<double>;
