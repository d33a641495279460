// common.cuh -- shared device/host helpers for the fdfd_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <complex>
#include <string>
#include <vector>
#include "../../include/fdfd_b200.h"

// physical constants, bit-identical to src/types.jl:6-9 of the reference
static constexpr double kEps0 = 8.85418782e-12;
static constexpr double kMu0 = 1.25663706e-6;

// ---- complex number on device: cplx<double> == double2 layout, cplx<float> == float2 layout
template <typename T> struct __align__(2 * sizeof(T)) cplx {
  T x, y;
  __host__ __device__ cplx() {}
  __host__ __device__ cplx(T re, T im = T(0)) : x(re), y(im) {}
  template <typename U> __host__ __device__ explicit cplx(const cplx<U>& o) : x(T(o.x)), y(T(o.y)) {}
};
using c128 = cplx<double>;
using c64 = cplx<float>;

template <typename T> __host__ __device__ __forceinline__ cplx<T> operator+(cplx<T> a, cplx<T> b) { return cplx<T>(a.x + b.x, a.y + b.y); }
template <typename T> __host__ __device__ __forceinline__ cplx<T> operator-(cplx<T> a, cplx<T> b) { return cplx<T>(a.x - b.x, a.y - b.y); }
template <typename T> __host__ __device__ __forceinline__ cplx<T> operator-(cplx<T> a) { return cplx<T>(-a.x, -a.y); }
template <typename T> __host__ __device__ __forceinline__ cplx<T> operator*(cplx<T> a, cplx<T> b) {
  return cplx<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T> __host__ __device__ __forceinline__ cplx<T> operator*(T s, cplx<T> a) { return cplx<T>(s * a.x, s * a.y); }
template <typename T> __host__ __device__ __forceinline__ cplx<T> operator*(cplx<T> a, T s) { return cplx<T>(s * a.x, s * a.y); }
template <typename T> __host__ __device__ __forceinline__ cplx<T>& operator+=(cplx<T>& a, cplx<T> b) { a.x += b.x; a.y += b.y; return a; }
template <typename T> __host__ __device__ __forceinline__ cplx<T>& operator-=(cplx<T>& a, cplx<T> b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename T> __host__ __device__ __forceinline__ cplx<T> conj(cplx<T> a) { return cplx<T>(a.x, -a.y); }
template <typename T> __host__ __device__ __forceinline__ T norm2(cplx<T> a) { return a.x * a.x + a.y * a.y; }
// a += b*c  (4 FMAs)
template <typename T> __host__ __device__ __forceinline__ void cfma(cplx<T>& a, cplx<T> b, cplx<T> c) {
  a.x = fma(b.x, c.x, a.x); a.x = fma(-b.y, c.y, a.x);
  a.y = fma(b.x, c.y, a.y); a.y = fma(b.y, c.x, a.y);
}
template <typename T> __host__ __device__ __forceinline__ cplx<T> crecip(cplx<T> a) {
  T d = T(1) / (a.x * a.x + a.y * a.y);
  return cplx<T>(a.x * d, -a.y * d);
}
template <typename T> __host__ __device__ __forceinline__ cplx<T> cdiv(cplx<T> a, cplx<T> b) { return a * crecip(b); }
// conj(a)*b
template <typename T> __host__ __device__ __forceinline__ cplx<T> cmulc(cplx<T> a, cplx<T> b) {
  return cplx<T>(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}

static inline c128 to_c128(std::complex<double> z) { return c128(z.real(), z.imag()); }

// ---- error plumbing ------------------------------------------------------------------
void fdfd_set_error(fdfd_ctx* ctx, const char* fmt, ...);

#define CUDA_TRY(ctx, expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      fdfd_set_error(ctx, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return FDFD_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

#define FDFD_TRY(expr)          \
  do {                          \
    int _s = (expr);            \
    if (_s != FDFD_OK) return _s; \
  } while (0)

#define ARG_CHECK(ctx, cond, msg)                                   \
  do {                                                              \
    if (!(cond)) {                                                  \
      fdfd_set_error(ctx, "%s: %s", __func__, msg);                 \
      return FDFD_ERR_ARG;                                          \
    }                                                               \
  } while (0)

// ---- context ---------------------------------------------------------------------------
struct fdfd_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 148;
  int64_t launches = 0;
  std::string err;
  std::vector<void*> scratch;  // freed at destroy
};

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with an unlimited release threshold: a buffer
// freed by one solve is handed to the next one without a trip to the driver.  Plain cudaMalloc / cudaFree synchronise the
// whole device, so with several frequencies in flight (fdfd_solve_driven, one worker stream each) every per-frequency setup and
// tear-down used to stall behind the other workers' kernels (round 1: setup_ms bursts of 0.6-2.9 s).  Allocation and release go
// through one allocator stream per device that never runs a kernel; the allocation is synchronised on that stream only, so
// the memory is valid on every stream when alloc() returns.  Release is safe because every owner synchronises the stream it
// used the buffer on before the buffer's destructor runs (every C-ABI call syncs before it returns).
cudaError_t fdfd_dev_alloc(void** p, size_t bytes);
void fdfd_dev_free(void* p);
size_t fdfd_dev_mem_available();   // free device memory + what the pool holds cached (bytes, current device)

// RAII device buffer (sizes here are few and large)
template <typename T> struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  bool own = true;  // false: p aliases memory owned elsewhere
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), own(o.own) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept { release(); p = o.p; n = o.n; own = o.own; o.p = nullptr; o.n = 0; return *this; }
  ~DevBuf() { release(); }
  void release() { if (p && own) fdfd_dev_free(p); p = nullptr; n = 0; own = true; }
  void alias(T* q, size_t count) { release(); p = q; n = count; own = false; }
  cudaError_t alloc(size_t count) {
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = fdfd_dev_alloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
};

// copy count elements from a caller pointer (host or device) into device memory, async on stream
bool fdfd_is_device_ptr(const void* p);
int fdfd_copy_in(fdfd_ctx* ctx, void* dst_dev, const void* src_any, size_t bytes);
int fdfd_copy_out(fdfd_ctx* ctx, void* dst_any, const void* src_dev, size_t bytes);

#define KLAUNCH(ctx) ((ctx)->launches++)

// ---- 1-D coefficient set of one operator (host side), see pml.cu --------------------------
// (A u)[ix,iy] = gxW*cxm[ix](u[ix-1]-u) + gxE*cxp[ix](u[ix+1]-u) + gyS*cym[iy](u[iy-1]-u) + gyN*cyp[iy](u[iy+1]-u) + m u
struct Coef1D {
  std::vector<std::complex<double>> cxm, cxp, cym, cyp;
};
void host_sfactor(const fdfd_grid_t& g, int dir, int fwd, double omega, std::vector<std::complex<double>>& s);
// continuous s-profile sampled at (1-based, fractional) position p along dir; equals host_sfactor at
// p=i (backward) / p=i+0.5 (forward)
std::complex<double> host_sprofile(const fdfd_grid_t& g, int dir, double omega, double p);
// fine-level coefficients from the exact reference s-factors.  scale = 1/mu0 (TM) or 1 (TE/eigen)
void host_coef_fine(const fdfd_grid_t& g, double omega, int ordering, double scale, Coef1D& c);
// rediscretised coefficients on a coarse level whose points sit at fine 1-based positions 1 + I*stride
void host_coef_level(const fdfd_grid_t& g, double omega, int ordering, double scale, int64_t stride,
                     int64_t nxl, int64_t nyl, Coef1D& c);

static inline double grid_dx(const fdfd_grid_t& g) { return (g.x1 - g.x0) / (double)g.Nx; }
static inline double grid_dy(const fdfd_grid_t& g) { return (g.y1 - g.y0) / (double)g.Ny; }
int check_grid(fdfd_ctx* ctx, const fdfd_grid_t* g);
