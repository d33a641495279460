// ctx.cu -- context, error reporting, host<->device staging, PML s-factors / 1-D coefficients.
#include "common.cuh"
#include <cstdarg>
#include <cmath>
#include <mutex>

static std::string g_last_error;
static std::mutex g_err_mu;

void fdfd_set_error(fdfd_ctx* ctx, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_last_error = buf;
}

extern "C" int fdfd_abi_version(void) { return FDFD_B200_ABI_VERSION; }

extern "C" const char* fdfd_last_error(fdfd_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  return g_last_error.c_str();
}

extern "C" int64_t fdfd_launch_count(fdfd_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- pooled device memory (see common.cuh) -------------------------------------------------------------------------------
namespace {
struct AllocState { cudaStream_t stream = nullptr; bool pooled = false; bool init = false; };
AllocState g_alloc[64];
std::mutex g_alloc_mu;
AllocState& alloc_state(int dev) {
  std::lock_guard<std::mutex> lk(g_alloc_mu);
  AllocState& A = g_alloc[dev & 63];
  if (!A.init) {
    A.init = true;
    const char* e = getenv("FDFD_NO_POOL");   // diagnostics: plain cudaMalloc / cudaFree
    int supported = 0;
    if (!(e && atoi(e) != 0) && cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev) == cudaSuccess && supported) {
      cudaMemPool_t pool = nullptr;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && cudaStreamCreateWithFlags(&A.stream, cudaStreamNonBlocking) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        A.pooled = true;
      }
    }
    cudaGetLastError();
  }
  return A;
}
}  // namespace

cudaError_t fdfd_dev_alloc(void** p, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  AllocState& A = alloc_state(dev);
  if (!A.pooled) return cudaMalloc(p, bytes);
  cudaError_t e = cudaMallocAsync(p, bytes, A.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(A.stream);
  return e;
}

size_t fdfd_dev_mem_available() {
  size_t fr = 0, tot = 0;
  if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return 0; }
  int dev = 0;
  cudaGetDevice(&dev);
  if (alloc_state(dev).pooled) {
    cudaMemPool_t pool = nullptr;
    uint64_t reserved = 0, used = 0;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used) fr += (size_t)(reserved - used);
    cudaGetLastError();
  }
  return fr;
}

void fdfd_dev_free(void* p) {
  if (!p) return;
  cudaPointerAttributes a;
  int dev = 0;
  if (cudaPointerGetAttributes(&a, p) == cudaSuccess) dev = a.device; else { cudaGetLastError(); cudaGetDevice(&dev); }
  AllocState& A = alloc_state(dev);
  if (!A.pooled) { cudaFree(p); return; }
  if (cudaFreeAsync(p, A.stream) != cudaSuccess) { cudaGetLastError(); cudaFree(p); }
}

extern "C" int fdfd_ctx_create(int device, void* stream, fdfd_ctx** out) {
  if (!out) { fdfd_set_error(nullptr, "fdfd_ctx_create: out is NULL"); return FDFD_ERR_ARG; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    // no CPU fallback by design: fail loudly
    fdfd_set_error(nullptr, "fdfd_ctx_create: no CUDA device (%s); this library has no CPU path",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return FDFD_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { fdfd_set_error(nullptr, "fdfd_ctx_create: bad device %d of %d", device, ndev); return FDFD_ERR_ARG; }
  CUDA_TRY(nullptr, cudaSetDevice(device));
  fdfd_ctx* ctx = new fdfd_ctx();
  ctx->device = device;
  if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
  else {
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { fdfd_set_error(nullptr, "cudaStreamCreate: %s", cudaGetErrorString(e)); delete ctx; return FDFD_ERR_CUDA; }
    ctx->own_stream = true;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->num_sms = prop.multiProcessorCount;
  *out = ctx;
  return FDFD_OK;
}

extern "C" void fdfd_ctx_destroy(fdfd_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" void fdfd_default_opts(fdfd_solve_opts_t* o) {
  if (!o) return;
  o->solver = FDFD_SOLVER_AUTO;
  o->precond = FDFD_PRECOND_MG;
  o->tol = 1e-10;
  o->maxit = 20000;
  o->mg_precision = FDFD_MG_F32;
  o->mg_cycle = FDFD_CYCLE_W;
  o->mg_wdepth = 3;
  o->mg_nu1 = 1; o->mg_nu2 = 1;
  o->mg_coarse_sweeps = 2;
  o->mg_beta = 0.5;
  o->mg_wjac = 0.7;
  o->mg_wline = 0.6;
  o->check_every = 8;
  o->verbose = 0;
  o->mg_shift_growth = 0.0;
  o->mg_max_levels = 32;
  o->use_graph = 1;
  o->concurrency = 4;
  o->ml_spec = 0;
}

bool fdfd_is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int fdfd_copy_in(fdfd_ctx* ctx, void* dst_dev, const void* src_any, size_t bytes) {
  if (bytes == 0) return FDFD_OK;
  cudaMemcpyKind k = fdfd_is_device_ptr(src_any) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CUDA_TRY(ctx, cudaMemcpyAsync(dst_dev, src_any, bytes, k, ctx->stream));
  // pageable host memory: the copy is staged synchronously by the runtime; pinned: truly async.
  // Either way the caller's buffer must stay valid until we sync, which every API call does before returning.
  return FDFD_OK;
}

int fdfd_copy_out(fdfd_ctx* ctx, void* dst_any, const void* src_dev, size_t bytes) {
  if (bytes == 0) return FDFD_OK;
  cudaMemcpyKind k = fdfd_is_device_ptr(dst_any) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  CUDA_TRY(ctx, cudaMemcpyAsync(dst_any, src_dev, bytes, k, ctx->stream));
  return FDFD_OK;
}

int check_grid(fdfd_ctx* ctx, const fdfd_grid_t* g) {
  ARG_CHECK(ctx, g != nullptr, "grid is NULL");
  ARG_CHECK(ctx, g->Nx >= 3 && g->Ny >= 3, "grid needs Nx,Ny >= 3 (periodic 5-point stencil)");
  ARG_CHECK(ctx, g->Npml_x >= 0 && g->Npml_y >= 0, "negative Npml");
  ARG_CHECK(ctx, 2 * g->Npml_x <= g->Nx && 2 * g->Npml_y <= g->Ny, "PML thicker than the grid");
  ARG_CHECK(ctx, g->x1 > g->x0 && g->y1 > g->y0 && g->L0 > 0, "degenerate bounds / L0");
  ARG_CHECK(ctx, g->Nx * g->Ny < (int64_t(1) << 40), "grid too large");
  return FDFD_OK;
}

// ---------------------------------------------------------------------------------------------
// PML s-factors: src/pml.jl:1-31.  sigma_max = -(m+1) lnR / (2 eta0 Tw), S(l) = 1 - i sigma(l)/(w eps0 L0),
// m = 3.5, lnR = -12 (not reachable from S_create, so fixed here too).
// ---------------------------------------------------------------------------------------------
static const double kPmlM = 3.5, kPmlLnR = -12.0;

static inline std::complex<double> s_of_depth(double l, double Tw, double omega, double L0) {
  const double eta0 = std::sqrt(kMu0 / kEps0);
  const double sigma_max = -(kPmlM + 1) * kPmlLnR / (2 * eta0 * Tw);
  const double sig = sigma_max * std::pow(l / Tw, kPmlM);
  return std::complex<double>(1.0, -sig / (omega * (kEps0 * L0)));
}

void host_sfactor(const fdfd_grid_t& g, int dir, int fwd, double omega, std::vector<std::complex<double>>& s) {
  const int64_t Nw = dir == 0 ? g.Nx : g.Ny;
  const int64_t Np = dir == 0 ? g.Npml_x : g.Npml_y;
  const double dw = dir == 0 ? grid_dx(g) : grid_dy(g);
  s.assign(Nw, std::complex<double>(1.0, 0.0));
  if (Np == 0) return;  // reference would form Tw = 0 (sigma_max = Inf) but never use it
  const double Tw = Np * dw;
  for (int64_t i = 1; i <= Nw; ++i) {  // 1-based like the reference
    if (fwd) {
      if (i <= Np) s[i - 1] = s_of_depth(dw * (Np - i + 0.5), Tw, omega, g.L0);
      else if (i > Nw - Np) s[i - 1] = s_of_depth(dw * (i - (Nw - Np) - 0.5), Tw, omega, g.L0);
    } else {
      if (i <= Np) s[i - 1] = s_of_depth(dw * (Np - i + 1), Tw, omega, g.L0);
      else if (i > Nw - Np) s[i - 1] = s_of_depth(dw * (i - (Nw - Np) - 1), Tw, omega, g.L0);
    }
  }
}

std::complex<double> host_sprofile(const fdfd_grid_t& g, int dir, double omega, double p) {
  const int64_t Nw = dir == 0 ? g.Nx : g.Ny;
  const int64_t Np = dir == 0 ? g.Npml_x : g.Npml_y;
  if (Np == 0) return std::complex<double>(1.0, 0.0);
  const double dw = dir == 0 ? grid_dx(g) : grid_dy(g);
  const double Tw = Np * dw;
  p = std::fmod(p - 1.0, (double)Nw);
  if (p < 0) p += Nw;
  p += 1.0;  // in [1, Nw+1)
  double depth = std::max(std::max((Np + 1) - p, p - (Nw - Np + 1)), 0.0) * dw;
  if (depth <= 0) return std::complex<double>(1.0, 0.0);
  return s_of_depth(depth, Tw, omega, g.L0);
}

static void coef_from_inv(const std::vector<std::complex<double>>& sf, const std::vector<std::complex<double>>& sb,
                          double a, double scale, int ordering,
                          std::vector<std::complex<double>>& cm, std::vector<std::complex<double>>& cp) {
  // sf, sb: INVERSE s-factors.  a = 1/dw.  Products ordered like the sparse products of the reference:
  // f.b: A[n,n-1] = (sf_n a)(scale)(sb_n a), A[n,n+1] = (sf_n a)(scale)(sb_{n+1} a)       (driven.jl:35)
  // b.f: A[n,n+1] = (sb_n a)(scale)(sf_n a), A[n,n-1] = (sb_n a)(scale)(sf_{n-1} a)       (modulation.jl:82)
  const int64_t n = (int64_t)sf.size();
  cm.resize(n); cp.resize(n);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t ip = (i + 1) % n, im = (i + n - 1) % n;
    if (ordering == FDFD_ORDER_FB) {
      cm[i] = (sf[i] * a) * scale * (sb[i] * a);
      cp[i] = (sf[i] * a) * scale * (sb[ip] * a);
    } else {
      cp[i] = (sb[i] * a) * scale * (sf[i] * a);
      cm[i] = (sb[i] * a) * scale * (sf[im] * a);
    }
  }
}

void host_coef_fine(const fdfd_grid_t& g, double omega, int ordering, double scale, Coef1D& c) {
  std::vector<std::complex<double>> sxf, sxb, syf, syb;
  host_sfactor(g, 0, 1, omega, sxf); host_sfactor(g, 0, 0, omega, sxb);
  host_sfactor(g, 1, 1, omega, syf); host_sfactor(g, 1, 0, omega, syb);
  for (auto* v : {&sxf, &sxb, &syf, &syb}) for (auto& z : *v) z = 1.0 / z;
  coef_from_inv(sxf, sxb, 1.0 / grid_dx(g), scale, ordering, c.cxm, c.cxp);
  coef_from_inv(syf, syb, 1.0 / grid_dy(g), scale, ordering, c.cym, c.cyp);
}

void host_coef_level(const fdfd_grid_t& g, double omega, int ordering, double scale, int64_t stride,
                     int64_t nxl, int64_t nyl, Coef1D& c) {
  auto one = [&](int dir, int64_t nl, double dw, std::vector<std::complex<double>>& cm, std::vector<std::complex<double>>& cp) {
    std::vector<std::complex<double>> sf(nl), sb(nl);
    // coarse s-factors = averages of the continuous profile over the coarse edge (backward) / dual cell (forward), so
    // every level keeps the PML's total complex stretched length  X = int S dp  (point sampling of the steep
    // (l/Tw)^3.5 profile does not, and the cycle diverges when the PML is thinner than a coarse cell).  With
    // stride 1 this is exactly the reference: sb_i = S(i), sf_i = S(i + 0.5)  (pml.jl:14-27).
    for (int64_t I = 0; I < nl; ++I) {
      const double pi = 1.0 + (double)(I * stride);
      std::complex<double> ab(0, 0), af(0, 0);
      for (int64_t j = 0; j < stride; ++j) {
        ab += host_sprofile(g, dir, omega, pi - 0.5 * (double)stride + (double)j + 0.5);
        af += host_sprofile(g, dir, omega, pi + (double)j + 0.5);
      }
      sb[I] = (double)stride / ab;
      sf[I] = (double)stride / af;
    }
    coef_from_inv(sf, sb, 1.0 / (dw * (double)stride), scale, ordering, cm, cp);
  };
  one(0, nxl, grid_dx(g), c.cxm, c.cxp);
  one(1, nyl, grid_dy(g), c.cym, c.cyp);
}

extern "C" int fdfd_sfactors(fdfd_ctx* ctx, const fdfd_grid_t* g, double omega,
                             fdfd_c128* sxf, fdfd_c128* sxb, fdfd_c128* syf, fdfd_c128* syb) {
  FDFD_TRY(check_grid(ctx, g));
  ARG_CHECK(ctx, sxf && sxb && syf && syb, "NULL output");
  ARG_CHECK(ctx, omega > 0, "omega must be > 0");
  std::vector<std::complex<double>> s;
  fdfd_c128* outs[4] = {sxf, sxb, syf, syb};
  for (int k = 0; k < 4; ++k) {
    host_sfactor(*g, k / 2, (k % 2) == 0, omega, s);
    std::memcpy(outs[k], s.data(), s.size() * sizeof(fdfd_c128));
  }
  return FDFD_OK;
}
