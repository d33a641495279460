// consumers.cu -- the callers either side of the hot path, kept on the device so that large sweeps do not
// have to move O(N) data over PCIe:
//   * fdfd_problem_flux_x : flux_surface_integral(field, center, width, x̂) for a TM field (src/flux.jl:37-47) evaluated
//     from the resident solution (one column: O(Ny) work, one double back instead of an 805 MB field at 4096^2)
//   * fdfd_rasterize      : setup_ϵᵣ!(d, shapes) (src/device.jl:47-61, `_compose_shapes!`) for boxes and cylinders:
//     the first shape (in list order) containing the pixel centre wins; replaces an O(N) host loop + KD-tree
#include "krylov.cuh"
#include "reduce.cuh"
#include <cmath>

namespace {

constexpr int kT = 256;

// sum_{iy in [j0,j1)} -0.5 Re( (Ez[xi,iy] + Ez[xi+1,iy])/2 * conj(Hy[xi,iy]) ),  Hy = k2 * (D_x Ez) at column xi
__global__ void __launch_bounds__(kT)
k_flux_x(int64_t Nx, int64_t Ny, const c128* __restrict__ ez, const c128* __restrict__ sx, double ax, int forward, c128 k2,
         int64_t xi, int64_t j0, int64_t j1, double* __restrict__ out) {
  double acc[1] = {0.0};
  const int64_t xm = xi == 0 ? Nx - 1 : xi - 1, xp = xi + 1 == Nx ? 0 : xi + 1;
  const c128 s = sx[xi];
  for (int64_t iy = j0 + threadIdx.x; iy < j1; iy += kT) {
    const c128 ec = ez[xi + Nx * iy], ee = ez[xp + Nx * iy];
    c128 dxu;
    if (forward) dxu = c128(s.x * -ax, s.y * -ax) * ec + c128(s.x * ax, s.y * ax) * ee;
    else dxu = c128(s.x * -ax, s.y * -ax) * ez[xm + Nx * iy] + c128(s.x * ax, s.y * ax) * ec;
    const c128 hy = k2 * dxu;
    const c128 ea((ec.x + ee.x) / 2, (ec.y + ee.y) / 2);
    acc[0] += -0.5 * (ea.x * hy.x + ea.y * hy.y);  // Re(ea * conj(hy))
  }
  block_reduce_store<kT, 1>(acc, out);
}

struct Shape { int kind; double cx, cy, a, b, eps_re, eps_im; };  // kind 0: box (a,b = full widths), 1: cylinder (a = radius)

__global__ void k_rasterize(int64_t Nx, int64_t Ny, double x0, double dx, double y0, double dy, int nshapes,
                            const Shape* __restrict__ shapes, c128* __restrict__ eps) {
  const int64_t N = Nx * Ny;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ix = n % Nx, iy = n / Nx;
    const double x = x0 + dx * (0.5 + (double)ix), y = y0 + dy * (0.5 + (double)iy);  // xc, yc (grid.jl:76-82)
    for (int k = 0; k < nshapes; ++k) {
      const Shape s = shapes[k];
      const bool in = s.kind == 0 ? (fabs(x - s.cx) <= s.a / 2 && fabs(y - s.cy) <= s.b / 2)
                                  : ((x - s.cx) * (x - s.cx) + (y - s.cy) * (y - s.cy) <= s.a * s.a);
      if (in) { eps[n] = c128(s.eps_re, s.eps_im); break; }
    }
  }
}

}  // namespace

extern "C" int fdfd_problem_flux_x(fdfd_problem* P, double center_x, double center_y, double width, int forward_h, double* flux) {
  if (!P) return FDFD_ERR_ARG;
  fdfd_ctx* ctx = P->ctx;
  ARG_CHECK(ctx, flux != nullptr, "flux is NULL");
  ARG_CHECK(ctx, P->op.pol == FDFD_TM, "flux_surface_integral is defined for TM fields only (flux.jl:48-55 is broken for TE)");
  const fdfd_grid_t& g = P->op.g;
  const double dx = grid_dx(g), dy = grid_dy(g);
  // x-index: the centre within dx/2 of center_x (lower index on a tie, like the notebook); y-range: centres in [cy-w, cy+w]
  int64_t xi = -1;
  for (int64_t i = 0; i < g.Nx; ++i) {
    const double xc = g.x0 + dx * (0.5 + (double)i);
    if (std::fabs(xc - center_x) <= dx / 2 * (1 + 1e-9)) { xi = i; break; }
  }
  ARG_CHECK(ctx, xi >= 0, "no x-centre within dx/2 of center_x");
  // flux.jl:40-42 indexes the column after xi: on the last column the reference throws, a periodic wrap to column 0 would be silent
  ARG_CHECK(ctx, xi + 1 < g.Nx, "the flux plane sits on the last grid column: column xi + 1 does not exist (flux.jl:40-42)");
  int64_t j0 = g.Ny, j1 = 0;
  for (int64_t j = 0; j < g.Ny; ++j) {
    const double yc = g.y0 + dy * (0.5 + (double)j);
    if (yc >= center_y - width && yc <= center_y + width) { j0 = std::min(j0, j); j1 = std::max(j1, j + 1); }
  }
  if (j1 <= j0) { *flux = 0.0; return FDFD_OK; }
  std::vector<std::complex<double>> sx;
  host_sfactor(g, 0, forward_h, P->op.omega_pml, sx);
  for (auto& z : sx) z = 1.0 / z;
  DevBuf<c128> dsx; DevBuf<double> dout;
  CUDA_TRY(ctx, dsx.alloc(g.Nx)); CUDA_TRY(ctx, dout.alloc(1));
  CUDA_TRY(ctx, cudaMemcpyAsync(dsx.p, sx.data(), g.Nx * sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
  const std::complex<double> I(0.0, 1.0);
  const std::complex<double> k2 = ((1.0 / I) / P->op.omega) / (kMu0 * g.L0);  // hy = 1/1im/ω/μ₀ * Dx ez
  k_flux_x<<<1, kT, 0, ctx->stream>>>(g.Nx, g.Ny, P->w.x.p, dsx.p, 1.0 / dx, forward_h, to_c128(k2), xi, j0, j1, dout.p);
  KLAUNCH(ctx);
  double h = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&h, dout.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *flux = h * dy;
  return FDFD_OK;
}

extern "C" int fdfd_rasterize(fdfd_ctx* ctx, const fdfd_grid_t* g, int nshapes, const double* shapes7, fdfd_c128* eps_r) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, nshapes >= 0 && (nshapes == 0 || shapes7) && eps_r, "bad arguments");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int64_t N = g->Nx * g->Ny;
  std::vector<Shape> hs(nshapes);
  for (int k = 0; k < nshapes; ++k) {
    const double* s = shapes7 + 7 * k;
    ARG_CHECK(ctx, s[0] == 0.0 || s[0] == 1.0, "shape kind must be 0 (box) or 1 (cylinder)");
    hs[k] = Shape{(int)s[0], s[1], s[2], s[3], s[4], s[5], s[6]};
  }
  DevBuf<Shape> ds; DevBuf<c128> de;
  CUDA_TRY(ctx, ds.alloc(std::max(1, nshapes))); CUDA_TRY(ctx, de.alloc(N));
  if (nshapes) CUDA_TRY(ctx, cudaMemcpyAsync(ds.p, hs.data(), nshapes * sizeof(Shape), cudaMemcpyHostToDevice, ctx->stream));
  FDFD_TRY(fdfd_copy_in(ctx, de.p, eps_r, N * sizeof(c128)));  // pixels outside every shape keep their value (device.jl:53)
  const int blocks = (int)std::min<int64_t>((N + 255) / 256, (int64_t)ctx->num_sms * 16);
  k_rasterize<<<blocks, 256, 0, ctx->stream>>>(g->Nx, g->Ny, g->x0, grid_dx(*g), g->y0, grid_dy(*g), nshapes, ds.p, de.p);
  KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  FDFD_TRY(fdfd_copy_out(ctx, eps_r, de.p, N * sizeof(c128)));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FDFD_OK;
}
