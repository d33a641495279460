// assemble.cu -- K2/K3: Yee derivative operators and the 5-point system matrix written by the GPU
// straight into CSR / CSC (no intermediate COO, no sparse-sparse products).
//
// Reference behaviour reproduced (file:line under /root/reference):
//   δ(w,s,g)            src/grid.jl:128-154   periodic differences, values (1/d)*(+-1.0)
//   S_create            src/pml.jl:33-63      row scaling by the INVERSE s-factors
//   TM system           src/solver/driven.jl:35      A = Dxf mu0^-1 Dxb + Dyf mu0^-1 Dyb + w^2 eps0 eps_r
//   TE system           src/solver/driven.jl:45      A = Dxf Teps_xi Dxb + Dyf Teps_yi Dyb + w^2 mu0 I
//   b.f ordering        src/solver/modulation.jl:82  A1 = Dxb/mu0 Dxf + Dyb/mu0 Dyf
// Column indices inside a row (row indices inside a column for CSC) are sorted ascending, like
// Julia's SparseMatrixCSC / SciPy canonical CSR, so index arrays compare bit-exactly.
#include "common.cuh"
#include "device_ops.cuh"

namespace {

struct Ent { int64_t j; c128 v; };

__device__ __forceinline__ void cswap(Ent& a, Ent& b) {
  if (b.j < a.j) { Ent t = a; a = b; b = t; }
}

// one thread per row (CSR) or per column (CSC) of a stretched/unstretched derivative operator.
// dir: 0=x,1=y; sgn=+1 forward / -1 backward; sinv: inverse s-factor 1-D array along dir (or null).
__global__ void k_assemble_deriv(int64_t Nx, int64_t Ny, int dir, int sgn, double a, const c128* __restrict__ sinv,
                                 int csc, int64_t base, int64_t* __restrict__ ptr, int64_t* __restrict__ ind,
                                 c128* __restrict__ val) {
  const int64_t N = Nx * Ny;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n <= N; n += (int64_t)gridDim.x * blockDim.x) {
    if (n == N) { ptr[N] = 2 * N + base; break; }
    ptr[n] = 2 * n + base;
    const int64_t ix = n % Nx, iy = n / Nx;
    const int64_t iw = dir == 0 ? ix : iy, Nw = dir == 0 ? Nx : Ny, stride = dir == 0 ? 1 : Nx;
    // neighbour along dir in direction +sgn (CSR: column of the off-diagonal entry of row n)
    //                      or direction -sgn (CSC: row whose off-diagonal entry lands in column n)
    const int step = csc ? -sgn : sgn;
    int64_t jw = iw + step;
    if (jw < 0) jw += Nw; else if (jw >= Nw) jw -= Nw;
    const int64_t m = n + (jw - iw) * stride;
    // the value belongs to the ROW: row n for CSR; for CSC diag entry row n, off-diag entry row m
    const c128 s_row_diag = sinv ? sinv[iw] : c128(1.0, 0.0);
    const c128 s_row_off = sinv ? (csc ? sinv[jw] : sinv[iw]) : c128(1.0, 0.0);
    // δ values are built as (1/d)*(+-1.0) (grid.jl:138-151): forward diag -1, off +1; backward diag +1, off -1
    const double vd = a * (double)(-sgn), vo = a * (double)(sgn);
    Ent e0{n, c128(s_row_diag.x * vd, s_row_diag.y * vd)};
    Ent e1{m, c128(s_row_off.x * vo, s_row_off.y * vo)};
    cswap(e0, e1);
    ind[2 * n] = e0.j + base; val[2 * n] = e0.v;
    ind[2 * n + 1] = e1.j + base; val[2 * n + 1] = e1.v;
  }
}

// value of the four off-diagonal couplings and the diagonal of row (ix,iy)
struct Row5 { c128 W, E, S, Nn, C; };

template <bool TE>
__device__ __forceinline__ Row5 row_values(const OpView<double>& op, int64_t ix, int64_t iy) {
  const int64_t Nx = op.nx, Ny = op.ny;
  const int64_t n = ix + Nx * iy;
  Row5 r;
  r.W = op.cxm[ix]; r.E = op.cxp[ix]; r.S = op.cym[iy]; r.Nn = op.cyp[iy];
  c128 m;
  if (TE) {
    const int64_t ixp = ix + 1 == Nx ? 0 : ix + 1, iyp = iy + 1 == Ny ? 0 : iy + 1;
    r.W = r.W * op.gx[n]; r.E = r.E * op.gx[ixp + Nx * iy];
    r.S = r.S * op.gy[n]; r.Nn = r.Nn * op.gy[ix + Nx * iyp];
    m = op.mass_const;
  } else {
    m = op.mass[n];
  }
  // summation order of the reference: (x-part + y-part) + mass
  c128 dx = -r.W - r.E, dy = -r.S - r.Nn;
  r.C = (dx + dy) + m;
  return r;
}

template <bool TE>
__global__ void k_assemble_system(OpView<double> op, int csc, int64_t base, int64_t* __restrict__ ptr,
                                  int64_t* __restrict__ ind, c128* __restrict__ val) {
  const int64_t Nx = op.nx, Ny = op.ny, N = Nx * Ny;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n <= N; n += (int64_t)gridDim.x * blockDim.x) {
    if (n == N) { ptr[N] = 5 * N + base; break; }
    ptr[n] = 5 * n + base;
    const int64_t ix = n % Nx, iy = n / Nx;
    const int64_t ixm = ix == 0 ? Nx - 1 : ix - 1, ixp = ix + 1 == Nx ? 0 : ix + 1;
    const int64_t iym = iy == 0 ? Ny - 1 : iy - 1, iyp = iy + 1 == Ny ? 0 : iy + 1;
    Ent e[5];
    if (!csc) {
      Row5 r = row_values<TE>(op, ix, iy);
      e[0] = {ix + Nx * iym, r.S}; e[1] = {ixm + Nx * iy, r.W}; e[2] = {n, r.C};
      e[3] = {ixp + Nx * iy, r.E}; e[4] = {ix + Nx * iyp, r.Nn};
    } else {
      // column n: entry from row r is A[r,n]; row (ix,iym) couples to n through its N coefficient, etc.
      e[0] = {ix + Nx * iym, row_values<TE>(op, ix, iym).Nn};
      e[1] = {ixm + Nx * iy, row_values<TE>(op, ixm, iy).E};
      e[2] = {n, row_values<TE>(op, ix, iy).C};
      e[3] = {ixp + Nx * iy, row_values<TE>(op, ixp, iy).W};
      e[4] = {ix + Nx * iyp, row_values<TE>(op, ix, iyp).S};
    }
    // 5-element sorting network (9 compare-exchanges)
    cswap(e[0], e[1]); cswap(e[3], e[4]); cswap(e[2], e[4]); cswap(e[2], e[3]); cswap(e[0], e[3]);
    cswap(e[0], e[2]); cswap(e[1], e[4]); cswap(e[1], e[3]); cswap(e[1], e[2]);
#pragma unroll
    for (int k = 0; k < 5; ++k) { ind[5 * n + k] = e[k].j + base; val[5 * n + k] = e[k].v; }
  }
}

}  // namespace

extern "C" int fdfd_assemble_derivative(fdfd_ctx* ctx, const fdfd_grid_t* g, double omega, int which, int stretched,
                                        int format, int index_base, int64_t* ptr, int64_t* ind, fdfd_c128* val) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, which >= 0 && which <= 3, "which must be FDFD_DXF..FDFD_DYB");
  ARG_CHECK(ctx, format == FDFD_CSR || format == FDFD_CSC, "bad format");
  ARG_CHECK(ctx, index_base == 0 || index_base == 1, "index_base must be 0 or 1");
  ARG_CHECK(ctx, ptr && ind && val, "NULL output");
  ARG_CHECK(ctx, !stretched || omega > 0, "omega must be > 0");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int64_t N = g->Nx * g->Ny;
  const int dir = which / 2, fwd = (which % 2) == 0;
  DevBuf<c128> dsinv; DevBuf<int64_t> dptr, dind; DevBuf<c128> dval;
  if (stretched) {
    std::vector<std::complex<double>> s;
    host_sfactor(*g, dir, fwd, omega, s);
    for (auto& z : s) z = 1.0 / z;  // `.^-1`, pml.jl:46-54
    CUDA_TRY(ctx, dsinv.alloc(s.size()));
    CUDA_TRY(ctx, cudaMemcpyAsync(dsinv.p, s.data(), s.size() * sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  CUDA_TRY(ctx, dptr.alloc(N + 1)); CUDA_TRY(ctx, dind.alloc(2 * N)); CUDA_TRY(ctx, dval.alloc(2 * N));
  const double a = 1.0 / (dir == 0 ? grid_dx(*g) : grid_dy(*g));
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((N + threads) / threads, (int64_t)ctx->num_sms * 16);
  k_assemble_deriv<<<blocks, threads, 0, ctx->stream>>>(g->Nx, g->Ny, dir, fwd ? 1 : -1, a, dsinv.p, format == FDFD_CSC,
                                                       (int64_t)index_base, dptr.p, dind.p, dval.p);
  KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  FDFD_TRY(fdfd_copy_out(ctx, ptr, dptr.p, (N + 1) * sizeof(int64_t)));
  FDFD_TRY(fdfd_copy_out(ctx, ind, dind.p, 2 * N * sizeof(int64_t)));
  FDFD_TRY(fdfd_copy_out(ctx, val, dval.p, 2 * N * sizeof(c128)));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FDFD_OK;
}

extern "C" int fdfd_assemble_system(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, int ordering, double omega,
                                    const fdfd_c128* eps_r, int format, int index_base, int64_t* ptr, int64_t* ind,
                                    fdfd_c128* val) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, pol == FDFD_TM || pol == FDFD_TE, "pol must be FDFD_TM or FDFD_TE");
  ARG_CHECK(ctx, ordering == FDFD_ORDER_FB || ordering == FDFD_ORDER_BF, "bad ordering");
  ARG_CHECK(ctx, format == FDFD_CSR || format == FDFD_CSC, "bad format");
  ARG_CHECK(ctx, index_base == 0 || index_base == 1, "index_base must be 0 or 1");
  ARG_CHECK(ctx, eps_r && ptr && ind && val, "NULL argument");
  ARG_CHECK(ctx, omega > 0, "omega must be > 0");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int64_t N = g->Nx * g->Ny;
  FineOp op;
  FDFD_TRY(op.build(ctx, *g, pol, ordering, omega, eps_r));
  DevBuf<int64_t> dptr, dind; DevBuf<c128> dval;
  CUDA_TRY(ctx, dptr.alloc(N + 1)); CUDA_TRY(ctx, dind.alloc(5 * N)); CUDA_TRY(ctx, dval.alloc(5 * N));
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((N + threads) / threads, (int64_t)ctx->num_sms * 16);
  if (pol == FDFD_TE)
    k_assemble_system<true><<<blocks, threads, 0, ctx->stream>>>(op.view(), format == FDFD_CSC, (int64_t)index_base, dptr.p, dind.p, dval.p);
  else
    k_assemble_system<false><<<blocks, threads, 0, ctx->stream>>>(op.view(), format == FDFD_CSC, (int64_t)index_base, dptr.p, dind.p, dval.p);
  KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  FDFD_TRY(fdfd_copy_out(ctx, ptr, dptr.p, (N + 1) * sizeof(int64_t)));
  FDFD_TRY(fdfd_copy_out(ctx, ind, dind.p, 5 * N * sizeof(int64_t)));
  FDFD_TRY(fdfd_copy_out(ctx, val, dval.p, 5 * N * sizeof(c128)));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FDFD_OK;
}
