// mlkrylov.cu -- multilevel Krylov solver (opt-in, FDFD_SOLVER_MLKRYLOV): the "deflation on top of the shifted-Laplacian
// preconditioner" of DESIGN.md §8 in the form the CPU prototypes (tools/twolevel_prototype.py, tools/multilevel_prototype.py)
// found to work with the library's REAL components.  Stands where dolinearsolve's `lu(A)\b` stands in the reference
// (src/solver/solver.jl:29-35); everything here is new work.
//
// Method (Erlangga & Nabben 2008, Sheikh et al. 2016, with two simplifications measured in the prototypes):
//   level l runs a flexible GMRES on A_l, right-preconditioned by the two-level operator (ADEF-1)
//       T_l v = q + M_l^-1 (v - A_l q),        q = Z_l  A_{l+1}^-1  (Z_l^T v / 4),
//   * M_l^-1 is ONE multigrid cycle of the existing shifted-Laplacian hierarchy started on level l (fp32),
//   * A_{l+1}^-1 is a FIXED number of iterations of the same method one level down (so the whole inner solve is a fixed
//     launch sequence with no host synchronisation: it is captured once and replayed as ONE CUDA graph),
//   * A_l (l >= 1) is the multigrid hierarchy's level operator with the complex shift removed (rediscretised, 5-point;
//     the Galerkin product Z^T A Z is a 9-point operator with 2-D coefficients and measured no better: 31 vs 32 outer
//     iterations at 512^2), so every level reuses the fp64 stencil kernel k_apply,
//   * Z_l is plain bilinear interpolation between the vertex-centred grids (coarse I <-> fine 2I), Z^T A Z ~ 4 A_{l+1},
//   * the last level is preconditioned by its multigrid cycle alone.
// Level 0 works on the reference operator itself in fp64, checks the TRUE residual ||b - A x|| / ||b|| and restarts.
// All Gram-Schmidt coefficients, the Hessenberg matrix and the small least-squares solve stay on the device.
//
// CPU prototype numbers (bench map, multigrid cycles started per level to reach 1e-10, tools/multilevel_prototype.py):
//   1024^2: default solver 474 fine cycles; steps (6,12): 24 fine + 144 level-1 + 1728 level-2;  (6,8): 26 + 156 + 1248.
//   2048^2: default solver 960 fine cycles; steps (6,8): 47 + 282 + 2256;  (8,12): 32 + 256 + 3072  (DESIGN.md 5b).
//
// Measured on a B200 (round 2, tools/gpu_mlkrylov.py, tools/gpu_r2_exp*.py; bench map, one solve, single stream):
//   1024^2: 27 outer iterations (prototype 26) / 268 ms against 216 BiCGSTAB iterations / 257 ms;  2048^2: 49 / 672 ms against 437 / 720 ms;
//   4096^2, F cycles, steps (6,6): 105 / 3.3-4.1 s against 1230 (W depth 3) / 5.0 s -- hence FDFD_SOLVER_AUTO's crossover at 2^22 points.
// GPU tests: tests/test_gpu_mlkrylov.py.
#include "krylov.cuh"
#include "reduce.cuh"
#include <algorithm>
#include <cmath>
#include <memory>

namespace {

constexpr int kT = 256;
constexpr int kMaxK = 128;  // largest FGMRES basis handled by the one-thread least-squares kernel

// ---- vector kernels (grid-stride) ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) k_ml_dot(int64_t N, const c128* __restrict__ a, const c128* __restrict__ b, double* __restrict__ partials) {
  double acc[2] = {0, 0};
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < N; i += (int64_t)gridDim.x * kT) {
    const c128 q = cmulc(a[i], b[i]);   // conj(a) b
    acc[0] += q.x; acc[1] += q.y;
  }
  block_reduce_store<kT, 2>(acc, partials + (size_t)blockIdx.x * 2);
}
// mode 0: *out = sum ; mode 1: *out = (sqrt(max(Re sum, 0)), 0)
__global__ void k_ml_dot_final(const double* __restrict__ partials, int nb, c128* __restrict__ out, int mode) {
  double res[2];
  final_reduce<kT, 2>(partials, nb, res);
  if (threadIdx.x == 0) *out = mode == 1 ? c128(sqrt(fmax(res[0], 0.0)), 0.0) : c128(res[0], res[1]);
}
__global__ void k_ml_axpy_neg(int64_t N, const c128* __restrict__ coef, const c128* __restrict__ v, c128* __restrict__ w) {
  const c128 c = *coef;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) w[i] -= c * v[i];
}
// out = w / Re(*nrm)  (zero vector when the norm is zero: a zero right-hand side gives a zero solution)
__global__ void k_ml_scale_inv(int64_t N, const c128* __restrict__ nrm, const c128* __restrict__ w, c128* __restrict__ out) {
  const double n = nrm->x, s = n > 0.0 ? 1.0 / n : 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = c128(w[i].x * s, w[i].y * s);
}
// x (+)= sum_j y[j] Z_j,  j < k
__global__ void k_ml_combine(int64_t N, int k, const c128* __restrict__ y, const c128* const* __restrict__ Z, c128* __restrict__ x, int accumulate) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    c128 s = accumulate ? x[i] : c128(0.0, 0.0);
    for (int j = 0; j < k; ++j) cfma(s, y[j], Z[j][i]);
    x[i] = s;
  }
}
// ---- fused classical Gram-Schmidt pass of the inner levels (measured in the prototype: same outer iteration count as modified
// Gram-Schmidt, tools/multilevel_prototype.py CGS=1): all <V_v, w> of a chunk in one pass over w, then one pass subtracting them
constexpr int kMD = 8;   // basis vectors per launch
struct VecPtrs { const c128* p[kMD]; };
__global__ void __launch_bounds__(kT) k_ml_multidot(int64_t N, int nv, VecPtrs V, const c128* __restrict__ w, double* __restrict__ partials) {
  double acc[2 * kMD];
#pragma unroll
  for (int v = 0; v < 2 * kMD; ++v) acc[v] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < N; i += (int64_t)gridDim.x * kT) {
    const c128 wi = w[i];
#pragma unroll
    for (int v = 0; v < kMD; ++v) {
      if (v < nv) { const c128 q = cmulc(V.p[v][i], wi); acc[2 * v] += q.x; acc[2 * v + 1] += q.y; }
    }
  }
  block_reduce_store<kT, 2 * kMD>(acc, partials + (size_t)blockIdx.x * 2 * kMD);
}
// one CTA per dot product: out[v] = sum over the nb block partials (fixed order)
__global__ void __launch_bounds__(kT) k_ml_multidot_final(const double* __restrict__ partials, int nb, c128* __restrict__ out) {
  const int v = blockIdx.x;
  double acc[2] = {0.0, 0.0};
  for (int b = threadIdx.x; b < nb; b += kT) { acc[0] += partials[(size_t)b * 2 * kMD + 2 * v]; acc[1] += partials[(size_t)b * 2 * kMD + 2 * v + 1]; }
  block_reduce_store<kT, 2>(acc, reinterpret_cast<double*>(out + v));
}
// w -= sum_v coef[v] V_v
__global__ void k_ml_multiaxpy(int64_t N, int nv, const c128* __restrict__ coef, VecPtrs V, c128* __restrict__ w) {
  c128 c[kMD];
#pragma unroll
  for (int v = 0; v < kMD; ++v) c[v] = v < nv ? coef[v] : c128(0.0, 0.0);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    c128 s = w[i];
#pragma unroll
    for (int v = 0; v < kMD; ++v) {
      if (v < nv) { const c128 t = c[v] * V.p[v][i]; s -= t; }
    }
    w[i] = s;
  }
}
// r = b - t, partial ||r||^2
__global__ void __launch_bounds__(kT) k_ml_resid(int64_t N, const c128* __restrict__ b, const c128* __restrict__ t, c128* __restrict__ r, double* __restrict__ partials) {
  double acc[2] = {0, 0};
  for (int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x; i < N; i += (int64_t)gridDim.x * kT) {
    const c128 ri = b[i] - t[i];
    r[i] = ri;
    acc[0] += norm2(ri);
  }
  block_reduce_store<kT, 2>(acc, partials + (size_t)blockIdx.x * 2);
}
// multigrid right-hand side (fp32, scaled): f = s * (v - t)   (t == nullptr: f = s * v)
__global__ void k_ml_to_mg(int64_t N, const c128* __restrict__ v, const c128* __restrict__ t, c64* __restrict__ f, double s) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    c128 d = v[i];
    if (t) d -= t[i];
    f[i] = c64(c128(s * d.x, s * d.y));
  }
}
// z = q + u   (q == nullptr: z = u), u the fp32 multigrid result
__global__ void k_ml_from_mg(int64_t N, const c128* __restrict__ q, const c64* __restrict__ u, c128* __restrict__ z) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    c128 s = c128(u[i]);
    if (q) s += q[i];
    z[i] = s;
  }
}

// ---- transfers between the vertex-centred grids (coarse I <-> fine 2I; nc = (n + 1) / 2) -------------------------------------
// 1-D prolongation weight pattern: fine even i <- coarse i/2 (1); fine odd i <- coarse (i-1)/2 and ((i+1)/2 mod nc) (1/2 each).
// candidates of the transposed gather of coarse I: fine 2I-1, 2I, 2I+1 with weights w[0..2]
__host__ __device__ __forceinline__ void rcand(int64_t I, int64_t n, int64_t idx[3], double w[3]) {
  idx[1] = 2 * I; w[1] = 1.0;
  int64_t il = 2 * I - 1;
  if (il < 0) { if (n & 1) { il = 0; w[0] = 0.0; } else { il = n - 1; w[0] = 0.5; } }   // periodic wrap: only an even grid has the odd point n-1
  else w[0] = 0.5;
  idx[0] = il;
  int64_t ir = 2 * I + 1;
  if (ir >= n) { ir = 0; w[2] = 0.0; } else w[2] = 0.5;
  idx[2] = ir;
}
// one coarse point of gc = scale * Z^T v
__host__ __device__ __forceinline__ c128 restrict_point(int64_t n, int64_t nx, int64_t ny, int64_t ncx, const c128* __restrict__ v, double scale) {
  const int64_t I = n % ncx, J = n / ncx;
  int64_t xi[3], yi[3]; double wx[3], wy[3];
  rcand(I, nx, xi, wx); rcand(J, ny, yi, wy);
  double sr = 0.0, si = 0.0;
  for (int b = 0; b < 3; ++b) {
    if (wy[b] == 0.0) continue;
    for (int a = 0; a < 3; ++a) {
      if (wx[a] == 0.0) continue;
      const c128 f = v[xi[a] + nx * yi[b]];
      const double w = wx[a] * wy[b];
      sr += w * f.x; si += w * f.y;
    }
  }
  return c128(scale * sr, scale * si);
}
// one fine point of q = Z y
__host__ __device__ __forceinline__ c128 prolong_point(int64_t n, int64_t nx, int64_t ncx, int64_t ncy, const c128* __restrict__ y) {
  const int64_t ix = n % nx, iy = n / nx;
  const int64_t I0 = ix >> 1, J0 = iy >> 1;
  const bool ox = ix & 1, oy = iy & 1;
  const int64_t I1 = ox ? (I0 + 1) % ncx : I0, J1 = oy ? (J0 + 1) % ncy : J0;
  const c128 a = y[I0 + ncx * J0];
  if (ox && oy) { const c128 b = y[I1 + ncx * J0], c = y[I0 + ncx * J1], d = y[I1 + ncx * J1]; return c128(0.25 * (a.x + b.x + c.x + d.x), 0.25 * (a.y + b.y + c.y + d.y)); }
  if (ox) { const c128 b = y[I1 + ncx * J0]; return c128(0.5 * (a.x + b.x), 0.5 * (a.y + b.y)); }
  if (oy) { const c128 c = y[I0 + ncx * J1]; return c128(0.5 * (a.x + c.x), 0.5 * (a.y + c.y)); }
  return a;
}
__global__ void k_ml_restrict(int64_t nx, int64_t ny, int64_t ncx, int64_t ncy, const c128* __restrict__ v, c128* __restrict__ gc, double scale) {
  const int64_t Nc = ncx * ncy;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < Nc; n += (int64_t)gridDim.x * blockDim.x) gc[n] = restrict_point(n, nx, ny, ncx, v, scale);
}
__global__ void k_ml_prolong(int64_t nx, int64_t ny, int64_t ncx, int64_t ncy, const c128* __restrict__ y, c128* __restrict__ q) {
  const int64_t N = nx * ny;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) q[n] = prolong_point(n, nx, ncx, ncy, y);
}

// ---- the small least-squares problem  min || beta e1 - H y ||,  H (k+1) x k upper Hessenberg (column major, ld), one thread ----
__host__ __device__ __forceinline__ double cabs_d(c128 a) { return hypot(a.x, a.y); }
// scratch (global, per level): R packed upper triangle [k (k + 1) / 2] | sn [k] | g [k + 1] | cs [k] (real parts)
__host__ __device__ void lsq_core(int k, int ld, const c128* __restrict__ H, const c128* __restrict__ beta, c128* __restrict__ y, c128* __restrict__ res,
                                  c128* __restrict__ scratch, int kcap) {
  // R is built column by column with Givens rotations G_i = [c s; -conj(s) c], c real
  c128* R = scratch;                                   // column j starts at j (j + 1) / 2
  c128* sn = scratch + (size_t)kcap * (kcap + 1) / 2;
  c128* g = sn + kcap;
  c128* cs = g + kcap + 1;
  g[0] = c128(beta->x, 0.0);
  c128 col[kMaxK + 1];
  for (int j = 0; j < k; ++j) {
    for (int i = 0; i <= j + 1; ++i) col[i] = H[(size_t)j * ld + i];
    for (int i = 0; i < j; ++i) {
      const c128 a = col[i], b = col[i + 1];
      const double c = cs[i].x;
      col[i] = c128(c * a.x, c * a.y) + sn[i] * b;
      col[i + 1] = c128(c * b.x, c * b.y) - conj(sn[i]) * a;
    }
    const c128 a = col[j], b = col[j + 1];
    const double na = cabs_d(a), nb = cabs_d(b), d = hypot(na, nb);
    double c = 1.0; c128 s(0.0, 0.0);
    if (d > 0.0) {
      if (na == 0.0) { c = 0.0; s = c128(1.0, 0.0); }
      else { c = na / d; const c128 ph = c128(a.x / na, a.y / na); s = ph * conj(b); s = c128(s.x / d, s.y / d); }   // s = (a/|a|) conj(b) / d
    }
    cs[j] = c128(c, 0.0); sn[j] = s;
    col[j] = c128(c * a.x, c * a.y) + s * b;
    for (int i = 0; i <= j; ++i) R[j * (j + 1) / 2 + i] = col[i];
    g[j + 1] = -(conj(s) * g[j]);
    g[j] = c128(c * g[j].x, c * g[j].y);
  }
  for (int i = k - 1; i >= 0; --i) {   // back substitution
    c128 acc = g[i];
    for (int j = i + 1; j < k; ++j) acc -= R[j * (j + 1) / 2 + i] * y[j];
    const c128 rii = R[i * (i + 1) / 2 + i];
    y[i] = norm2(rii) > 0.0 ? cdiv(acc, rii) : c128(0.0, 0.0);
  }
  if (res) *res = c128(cabs_d(g[k]), 0.0);
}
__global__ void k_ml_lsq(int k, int ld, const c128* __restrict__ H, const c128* __restrict__ beta, c128* __restrict__ y, c128* __restrict__ res,
                         c128* __restrict__ scratch, int kcap) {
  if (threadIdx.x == 0 && blockIdx.x == 0) lsq_core(k, ld, H, beta, y, res, scratch, kcap);
}

}  // namespace

// one level of the method
struct MLLevel {
  int l = 0;                    // multigrid level this solver works on
  int k = 0;                    // FGMRES steps per solve (level 0: restart length)
  int64_t nx = 0, ny = 0, N = 0;
  int nb = 0;                   // blocks of the vector kernels
  FineOp* op = nullptr;         // fp64 operator A_l
  std::unique_ptr<FineOp> own;  // l >= 1: rediscretised level operator (shift removed)
  std::vector<DevBuf<c128>> V, Z;
  DevBuf<const c128*> Zptr;
  DevBuf<c128> w, q, t;         // w = A z / scratch, q = Z y, t = A q
  DevBuf<c128> rhs, x;          // l >= 1: restricted right-hand side and solution (level 0 uses the problem's b, x)
  DevBuf<c128> lsq;             // scratch of the least-squares kernel
  DevBuf<c128> H, y, sc;        // Hessenberg (k+1) x k, least-squares solution, scalars: sc[0] = beta, sc[1] = residual estimate
  DevBuf<double> parts;
};

struct MLKrylov {
  fdfd_problem* P = nullptr;
  std::vector<std::unique_ptr<MLLevel>> lev;
  std::map<std::vector<void*>, IterGraph> graphs;   // the level-1 solve, one graph per multigrid buffer-rotation state
  bool use_graph = false;
  bool fused_gs = true;                              // inner levels: fused classical Gram-Schmidt (FDFD_ML_MGS=1 selects the modified one)
  bool rel_w = true;                                 // W recursion depth of a cycle counted from the level it is started on (FDFD_ML_RELW=0: from level 0)
  bool l0_cgs = true;                                // level 0: fused classical Gram-Schmidt (one pass over w per 8 basis vectors) instead of the modified one
                                                     // (FDFD_ML_L0CGS=0); measured at 4096^2: same 176 outer iterations, 6.95 -> 4.64 s
  int cycle_kind = FDFD_CYCLE_F;                     // cycle of M_l^-1 (FDFD_ML_CYCLE): F on every level -- a W cycle truncated at an absolute depth
                                                     // degenerates to V on the inner levels; measured at 4096^2, steps (6,6): W2 210 outer / 6.0 s, F 105 / 3.3-4.1 s
  int64_t cycles[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // multigrid cycles started per level (diagnostics)
  ~MLKrylov() { for (auto& kv : graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec); }
};
void mlkrylov_free(MLKrylov* m) { delete m; }

namespace {

#define MALLOC(buf, cnt) do { if ((buf).alloc(cnt) != cudaSuccess) { cudaGetLastError(); fdfd_set_error(ctx, "multilevel Krylov: out of device memory allocating %zu elements", (size_t)(cnt)); return FDFD_ERR_ALLOC; } } while (0)

int ml_setup(fdfd_problem* P, MLKrylov& M) {
  fdfd_ctx* ctx = P->ctx;
  Multigrid<float>* mg = P->mgf;
  M.P = P;
  // spec: ml_spec bytes = k1 | k2 << 8 | k3 << 16 | restart << 24 ; 0 -> (6, 12, -, 96).  Levels beyond the hierarchy are dropped.
  uint32_t spec = (uint32_t)P->opts.ml_spec;
  if (const char* e = getenv("FDFD_ML_SPEC")) spec = (uint32_t)strtoul(e, nullptr, 0);   // diagnostics
  int ks[4] = {(int)(spec >> 24) & 0xff, (int)spec & 0xff, (int)(spec >> 8) & 0xff, (int)(spec >> 16) & 0xff};
  if ((spec & 0xffffff) == 0) { ks[1] = 6; ks[2] = 6; ks[3] = 0; }   // measured at 4096^2 (F cycles): (6,6) 105 outer / 3.3-4.1 s, (6,8) 89 / 5.1 s, (8,8) 70 / 3.9 s, (4,4) 228 / 5.3 s
  if (ks[0] == 0) ks[0] = 96;
  int nl = 1;
  while (nl < 4 && ks[nl] > 0 && nl < mg->levels()) ++nl;
  for (int l = 0; l < nl; ++l) ARG_CHECK(ctx, ks[l] >= 1 && ks[l] <= kMaxK, "multilevel Krylov: iteration counts must be in [1, 128]");
  for (int l = 0; l < nl; ++l) {
    M.lev.emplace_back(new MLLevel());
    MLLevel& L = *M.lev.back();
    L.l = l; L.k = ks[l];
    L.nx = mg->lv[l].nx; L.ny = mg->lv[l].ny; L.N = L.nx * L.ny;
    L.nb = vec_blocks_for(ctx, L.N);
    if (l == 0) L.op = &P->op;
    else {
      L.own.reset(new FineOp());
      FDFD_TRY(L.own->build_level(ctx, P->op.g, P->op.pol, L.nx, L.ny, mg->lv[l].hc, P->op.omega, mg->lv[l].eps.p));
      L.op = L.own.get();
      MALLOC(L.rhs, L.N); MALLOC(L.x, L.N);
    }
    L.V.resize(L.k + 1); L.Z.resize(L.k);
    // level 0 grows its basis on demand (a solve that converges in 25 iterations must not hold 97 fine vectors: four
    // concurrent 4096^2 solves would otherwise pin 104 GB); the inner levels are small and use every vector in every solve
    if (l == 0) MALLOC(L.V[0], L.N);
    else { for (auto& b : L.V) MALLOC(b, L.N); for (auto& b : L.Z) MALLOC(b, L.N); }
    std::vector<const c128*> zp(L.k);
    for (int j = 0; j < L.k; ++j) zp[j] = L.Z[j].p;   // level 0: nullptr until the vector exists
    MALLOC(L.Zptr, L.k);
    CUDA_TRY(ctx, cudaMemcpyAsync(L.Zptr.p, zp.data(), L.k * sizeof(const c128*), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // zp is a local
    MALLOC(L.w, L.N);
    if (l + 1 < nl) { MALLOC(L.q, L.N); MALLOC(L.t, L.N); }
    MALLOC(L.H, (size_t)(L.k + 1) * L.k); MALLOC(L.y, L.k); MALLOC(L.sc, 4);
    MALLOC(L.lsq, (size_t)L.k * (L.k + 1) / 2 + 3 * (size_t)L.k + 2);
    MALLOC(L.parts, (size_t)L.nb * 2 * kMD);
  }
  cudaStream_t st = ctx->stream;
  M.use_graph = P->opts.use_graph && st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
  if (const char* e = getenv("FDFD_ML_MGS")) M.fused_gs = atoi(e) == 0;   // diagnostics
  if (const char* e = getenv("FDFD_ML_L0CGS")) M.l0_cgs = atoi(e) != 0;
  if (const char* e = getenv("FDFD_ML_RELW")) M.rel_w = atoi(e) != 0;
  if (const char* e = getenv("FDFD_ML_CYCLE")) { const int v = atoi(e); if (v >= FDFD_CYCLE_V && v <= FDFD_CYCLE_W) M.cycle_kind = v; }
  return FDFD_OK;
}

// <a, b> -> *out (device), mode as k_ml_dot_final
int ml_dot(fdfd_ctx* ctx, MLLevel& L, const c128* a, const c128* b, c128* out, int mode) {
  k_ml_dot<<<L.nb, kT, 0, ctx->stream>>>(L.N, a, b, L.parts.p); KLAUNCH(ctx);
  k_ml_dot_final<<<1, kT, 0, ctx->stream>>>(L.parts.p, L.nb, out, mode); KLAUNCH(ctx);
  return FDFD_OK;
}

int ml_solve_level(MLKrylov& M, int li);

// z = T_l v
int ml_precond(MLKrylov& M, int li, const c128* v, c128* z) {
  fdfd_problem* P = M.P;
  fdfd_ctx* ctx = P->ctx;
  Multigrid<float>* mg = P->mgf;
  cudaStream_t st = ctx->stream;
  MLLevel& L = *M.lev[li];
  const bool last = li + 1 == (int)M.lev.size();
  const c128* q = nullptr;
  const c128* t = nullptr;
  if (!last) {
    MLLevel& C = *M.lev[li + 1];
    k_ml_restrict<<<C.nb, 256, 0, st>>>(L.nx, L.ny, C.nx, C.ny, v, C.rhs.p, 0.25); KLAUNCH(ctx);
    FDFD_TRY(ml_solve_level(M, li + 1));
    k_ml_prolong<<<L.nb, 256, 0, st>>>(L.nx, L.ny, C.nx, C.ny, C.x.p, L.q.p); KLAUNCH(ctx);
    DotSpec d0;
    FDFD_TRY(launch_apply(ctx, L.op->view(), L.op->pol == FDFD_TE, L.q.p, false, L.t.p, d0));
    q = L.q.p; t = L.t.p;
  }
  k_ml_to_mg<<<L.nb, 256, 0, st>>>(L.N, v, t, mg->lv[L.l].f.p, mg->rhs_scale); KLAUNCH(ctx);
  mg->wbase = M.rel_w ? L.l : 0;
  const int crc = mg->cycle(L.l, true, M.cycle_kind);
  mg->wbase = 0;
  FDFD_TRY(crc);
  M.cycles[L.l < 8 ? L.l : 7]++;
  k_ml_from_mg<<<L.nb, 256, 0, st>>>(L.N, q, mg->lv[L.l].u.p, z); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  return FDFD_OK;
}

// one Arnoldi step j of level li: Z_j = T V_j, w = A Z_j, orthogonalise (modified Gram-Schmidt), V_{j+1}
int ml_arnoldi_step(MLKrylov& M, int li, int j) {
  fdfd_ctx* ctx = M.P->ctx;
  cudaStream_t st = ctx->stream;
  MLLevel& L = *M.lev[li];
  FDFD_TRY(ml_precond(M, li, L.V[j].p, L.Z[j].p));
  DotSpec d0;
  FDFD_TRY(launch_apply(ctx, L.op->view(), L.op->pol == FDFD_TE, L.Z[j].p, false, L.w.p, d0));
  c128* Hj = L.H.p + (size_t)j * (L.k + 1);
  if ((li == 0 && !M.l0_cgs) || (li > 0 && !M.fused_gs)) {   // modified Gram-Schmidt
    for (int i = 0; i <= j; ++i) {
      FDFD_TRY(ml_dot(ctx, L, L.V[i].p, L.w.p, Hj + i, 0));
      k_ml_axpy_neg<<<L.nb, 256, 0, st>>>(L.N, Hj + i, L.V[i].p, L.w.p); KLAUNCH(ctx);
    }
  } else {                        // one classical pass in chunks of kMD vectors: every dot is taken against the same w
    for (int v0 = 0; v0 <= j; v0 += kMD) {
      const int nv = std::min(kMD, j + 1 - v0);
      VecPtrs vp;
      for (int v = 0; v < kMD; ++v) vp.p[v] = L.V[v0 + (v < nv ? v : 0)].p;
      k_ml_multidot<<<L.nb, kT, 0, st>>>(L.N, nv, vp, L.w.p, L.parts.p); KLAUNCH(ctx);
      k_ml_multidot_final<<<nv, kT, 0, st>>>(L.parts.p, L.nb, Hj + v0); KLAUNCH(ctx);
    }
    for (int v0 = 0; v0 <= j; v0 += kMD) {
      const int nv = std::min(kMD, j + 1 - v0);
      VecPtrs vp;
      for (int v = 0; v < kMD; ++v) vp.p[v] = L.V[v0 + (v < nv ? v : 0)].p;
      k_ml_multiaxpy<<<L.nb, 256, 0, st>>>(L.N, nv, Hj + v0, vp, L.w.p); KLAUNCH(ctx);
    }
  }
  FDFD_TRY(ml_dot(ctx, L, L.w.p, L.w.p, Hj + j + 1, 1));
  k_ml_scale_inv<<<L.nb, 256, 0, st>>>(L.N, Hj + j + 1, L.w.p, L.V[j + 1].p); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  return FDFD_OK;
}

// level li >= 1: x = (approximately) A^-1 rhs by exactly k FGMRES steps from a zero guess; no host synchronisation
int ml_solve_level_body(MLKrylov& M, int li) {
  fdfd_ctx* ctx = M.P->ctx;
  cudaStream_t st = ctx->stream;
  MLLevel& L = *M.lev[li];
  FDFD_TRY(ml_dot(ctx, L, L.rhs.p, L.rhs.p, L.sc.p, 1));                                       // beta = ||rhs||
  k_ml_scale_inv<<<L.nb, 256, 0, st>>>(L.N, L.sc.p, L.rhs.p, L.V[0].p); KLAUNCH(ctx);
  for (int j = 0; j < L.k; ++j) FDFD_TRY(ml_arnoldi_step(M, li, j));
  k_ml_lsq<<<1, 32, 0, st>>>(L.k, L.k + 1, L.H.p, L.sc.p, L.y.p, L.sc.p + 1, L.lsq.p, L.k); KLAUNCH(ctx);
  k_ml_combine<<<L.nb, 256, 0, st>>>(L.N, L.k, L.y.p, L.Zptr.p, L.x.p, 0); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  return FDFD_OK;
}

void ml_get_state(MLKrylov& M, std::vector<void*>& v) {
  v.clear();
  Multigrid<float>* mg = M.P->mgf;
  for (size_t l = 1; l < mg->lv.size(); ++l) { v.push_back(mg->lv[l].u.p); v.push_back(mg->lv[l].tmp.p); }
}
void ml_set_state(MLKrylov& M, const std::vector<void*>& v) {
  size_t i = 0;
  Multigrid<float>* mg = M.P->mgf;
  for (size_t l = 1; l < mg->lv.size(); ++l) { mg->lv[l].u.p = (c64*)v[i++]; mg->lv[l].tmp.p = (c64*)v[i++]; }
}

// the level-1 solve is a fixed launch sequence (thousands of small kernels): captured once per buffer-rotation state of
// the multigrid levels >= 1 and replayed as one CUDA graph; deeper levels are part of that graph
int ml_solve_level(MLKrylov& M, int li) {
  fdfd_ctx* ctx = M.P->ctx;
  cudaStream_t st = ctx->stream;
  if (li != 1 || !M.use_graph) return ml_solve_level_body(M, li);
  std::vector<void*> key;
  ml_get_state(M, key);
  auto it = M.graphs.find(key);
  if (it == M.graphs.end()) {
    const int64_t l0 = ctx->launches;
    int64_t cyc0[8]; std::copy(M.cycles, M.cycles + 8, cyc0);
    CUDA_TRY(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    const int rc = ml_solve_level_body(M, li);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != FDFD_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    CUDA_TRY(ctx, ce);
    IterGraph ig;
    ig.nlaunch = ctx->launches - l0;
    ctx->launches = l0;
    for (int q = 0; q < 8; ++q) { ig.post.push_back((void*)(intptr_t)(M.cycles[q] - cyc0[q])); M.cycles[q] = cyc0[q]; }   // cycles per replay, first 8 entries
    CUDA_TRY(ctx, cudaGraphInstantiate(&ig.exec, graph, 0));
    cudaGraphDestroy(graph);
    std::vector<void*> post;
    ml_get_state(M, post);
    ig.post.insert(ig.post.end(), post.begin(), post.end());
    it = M.graphs.emplace(key, ig).first;
  } else {
    ml_set_state(M, std::vector<void*>(it->second.post.begin() + 8, it->second.post.end()));
  }
  CUDA_TRY(ctx, cudaGraphLaunch(it->second.exec, st));
  ctx->launches += it->second.nlaunch;
  for (int q = 0; q < 8; ++q) M.cycles[q] += (int64_t)(intptr_t)it->second.post[q];
  return FDFD_OK;
}

}  // namespace

// level 0: restarted flexible GMRES on the reference operator, true residual at every restart
int krylov_multilevel(fdfd_problem* P, fdfd_info_t* info) {
  fdfd_ctx* ctx = P->ctx;
  ARG_CHECK(ctx, P->mgf != nullptr, "the multilevel Krylov solver needs the fp32 multigrid preconditioner");
  ARG_CHECK(ctx, P->mgf->levels() >= 2, "the multilevel Krylov solver needs a multigrid hierarchy of at least 2 levels");
  if (!P->ml) {
    P->ml = new MLKrylov();
    const int st = ml_setup(P, *P->ml);
    if (st != FDFD_OK) { mlkrylov_free(P->ml); P->ml = nullptr; return st; }
  }
  MLKrylov& M = *P->ml;
  MLLevel& L = *M.lev[0];
  cudaStream_t st = ctx->stream;
  const fdfd_solve_opts_t& o = P->opts;
  const int64_t N = L.N;
  const int* saved_done = P->mgf->done;
  P->mgf->done = nullptr;   // the cycles of this solver are never skipped by the BiCGSTAB convergence flag
  struct Restore { Multigrid<float>* mg; const int* d; ~Restore() { mg->done = d; } } restore{P->mgf, saved_done};
  std::fill(M.cycles, M.cycles + 8, 0);

  struct Events { cudaEvent_t e0 = nullptr, e1 = nullptr; ~Events() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); } } ev;
  CUDA_TRY(ctx, cudaEventCreate(&ev.e0)); CUDA_TRY(ctx, cudaEventCreate(&ev.e1));
  const int64_t launches0 = ctx->launches;
  CUDA_TRY(ctx, cudaEventRecord(ev.e0, st));

  c128* b = P->w.b.p; c128* x = P->w.x.p; c128* r = P->w.r.p; c128* t = P->w.t.p;
  c128 hs[2];
  auto fetch = [&](const c128* dev, int n) -> int {
    CUDA_TRY(ctx, cudaMemcpyAsync(hs, dev, n * sizeof(c128), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return FDFD_OK;
  };
  CUDA_TRY(ctx, cudaMemsetAsync(x, 0, N * sizeof(c128), st));
  FDFD_TRY(ml_dot(ctx, L, b, b, L.sc.p + 2, 1));
  FDFD_TRY(fetch(L.sc.p + 2, 1));
  const double bnorm = hs[0].x;
  int its = 0, restarts = 0, flag = FDFD_ERR_NOCONV;
  std::vector<double> hist(1, bnorm * bnorm);   // ||r_k||^2 (FGMRES estimate) per outer iteration, for fdfd_problem_get_history
  double rel = 0.0;
  if (bnorm == 0.0) { flag = FDFD_OK; }   // b = 0 -> x = 0
  else {
    CUDA_TRY(ctx, cudaMemcpyAsync(r, b, N * sizeof(c128), cudaMemcpyDeviceToDevice, st));
    double rnorm = bnorm, prev_rel = 1.0;
    int stalled = 0, kcap = L.k;   // kcap: restart length, shrunk if the device cannot hold the whole basis
    while (true) {
      // ---- one FGMRES cycle from the current residual r (||r|| = rnorm)
      const c128 hb(rnorm, 0.0);
      CUDA_TRY(ctx, cudaMemcpyAsync(L.sc.p, &hb, sizeof(c128), cudaMemcpyHostToDevice, st));
      k_ml_scale_inv<<<L.nb, 256, 0, st>>>(N, L.sc.p, r, L.V[0].p); KLAUNCH(ctx);
      CUDA_TRY(ctx, cudaStreamSynchronize(st));   // hb is a local
      int j = 0;
      for (; j < kcap && its < o.maxit; ) {
        if (!L.Z[j].p || !L.V[j + 1].p) {   // grow the basis; out of memory = restart here with what there is
          const bool ok = (L.Z[j].p || L.Z[j].alloc(N) == cudaSuccess) && (L.V[j + 1].p || L.V[j + 1].alloc(N) == cudaSuccess);
          if (!ok) {
            cudaGetLastError();
            if (j == 0) { fdfd_set_error(ctx, "multilevel Krylov: out of device memory for the first basis vectors"); return FDFD_ERR_ALLOC; }
            kcap = j;
            if (o.verbose) fprintf(stderr, "[fdfd_b200] multilevel Krylov: basis capped at %d vectors by device memory\n", kcap);
            break;
          }
          const c128* zp = L.Z[j].p;
          CUDA_TRY(ctx, cudaMemcpyAsync(L.Zptr.p + j, &zp, sizeof(zp), cudaMemcpyHostToDevice, st));
          CUDA_TRY(ctx, cudaStreamSynchronize(st));   // zp is a local
        }
        FDFD_TRY(ml_arnoldi_step(M, 0, j));
        ++j; ++its;
        k_ml_lsq<<<1, 32, 0, st>>>(j, L.k + 1, L.H.p, L.sc.p, L.y.p, L.sc.p + 1, L.lsq.p, L.k); KLAUNCH(ctx);
        FDFD_TRY(fetch(L.sc.p + 1, 1));
        const double est = hs[0].x / bnorm;
        hist.push_back(hs[0].x * hs[0].x);
        if (o.verbose) fprintf(stderr, "[fdfd_b200] multilevel Krylov it %d estimated relres %.3e\n", its, est);
        if (!std::isfinite(est)) { flag = FDFD_ERR_BREAKDOWN; break; }
        if (est <= o.tol) break;
      }
      if (flag == FDFD_ERR_BREAKDOWN) break;
      if (j > 0) { k_ml_combine<<<L.nb, 256, 0, st>>>(N, j, L.y.p, L.Zptr.p, x, 1); KLAUNCH(ctx); }
      // ---- true residual with the fp64 operator
      DotSpec d0;
      FDFD_TRY(launch_apply(ctx, P->op.view(), P->op.pol == FDFD_TE, x, false, t, d0));
      k_ml_resid<<<L.nb, kT, 0, st>>>(N, b, t, r, L.parts.p); KLAUNCH(ctx);
      k_ml_dot_final<<<1, kT, 0, st>>>(L.parts.p, L.nb, L.sc.p + 3, 1); KLAUNCH(ctx);
      FDFD_TRY(fetch(L.sc.p + 3, 1));
      rnorm = hs[0].x;
      rel = rnorm / bnorm;
      if (o.verbose) fprintf(stderr, "[fdfd_b200] multilevel Krylov restart %d: true relres %.3e after %d iterations\n", restarts, rel, its);
      if (std::isfinite(rel) && rel <= o.tol) { flag = FDFD_OK; break; }
      if (!std::isfinite(rel)) { flag = FDFD_ERR_BREAKDOWN; break; }
      if (its >= o.maxit) { flag = FDFD_ERR_NOCONV; break; }
      // stagnation: four restart cycles in a row that each gained less than a factor 2 -> give up instead of running to maxit
      stalled = rel > 0.5 * prev_rel ? stalled + 1 : 0;
      prev_rel = rel;
      if (stalled >= 4) { flag = FDFD_ERR_NOCONV; break; }
      ++restarts;
    }
  }
  cudaEventRecord(ev.e1, st);
  cudaEventSynchronize(ev.e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, ev.e0, ev.e1);
  {  // residual history in the layout of the BiCGSTAB loop (KrylovWork::hist, h_scal)
    const size_t cnt = std::min(hist.size(), P->w.hist.n);
    cudaMemcpyAsync(P->w.hist.p, hist.data(), cnt * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);   // hist is a local
    P->w.h_scal->iter = (int)cnt - 1; P->w.h_scal->bnorm2 = bnorm * bnorm;
  }
  info->iters = its;
  info->relres = rel;
  info->flag = flag;
  info->restarts = restarts;
  info->solve_ms = ms;
  info->launches = ctx->launches - launches0;
  if (o.verbose) fprintf(stderr, "[fdfd_b200] multilevel Krylov: multigrid cycles per level %lld %lld %lld %lld\n", (long long)M.cycles[0],
                         (long long)M.cycles[1], (long long)M.cycles[2], (long long)M.cycles[3]);
  return FDFD_OK;
}

// diagnostics: multigrid cycles started per level during the last multilevel solve
extern "C" int fdfd_problem_ml_cycles(fdfd_problem* P, int64_t* out4) {
  if (!P || !out4) return FDFD_ERR_ARG;
  for (int q = 0; q < 4; ++q) out4[q] = P->ml ? P->ml->cycles[q] : 0;
  return FDFD_OK;
}

// host-only test hooks (no GPU needed): the arithmetic cores of the kernels above run on the CPU so that the `-m "not gpu"`
// suite can check them -- the small least-squares solve against a dense solver, the transfers for adjointness
extern "C" int fdfd_debug_ml_lsq(int k, const fdfd_c128* H, double beta, fdfd_c128* y, double* resnorm) {
  if (k < 1 || k > kMaxK || !H || !y) return FDFD_ERR_ARG;
  std::vector<c128> scratch((size_t)k * (k + 1) / 2 + 3 * (size_t)k + 2), yy(k);
  const c128 b(beta, 0.0);
  c128 res(0.0, 0.0);
  lsq_core(k, k + 1, reinterpret_cast<const c128*>(H), &b, yy.data(), &res, scratch.data(), k);
  std::memcpy(y, yy.data(), sizeof(c128) * k);
  if (resnorm) *resnorm = res.x;
  return FDFD_OK;
}
// mode 0: out (ncx x ncy) = scale * Z^T in (nx x ny);  mode 1: out (nx x ny) = Z in (ncx x ncy);  nc = (n + 1) / 2
extern "C" int fdfd_debug_ml_transfer(int64_t nx, int64_t ny, int mode, double scale, const fdfd_c128* in, fdfd_c128* out) {
  if (nx < 2 || ny < 2 || !in || !out) return FDFD_ERR_ARG;
  const int64_t ncx = (nx + 1) / 2, ncy = (ny + 1) / 2;
  const c128* a = reinterpret_cast<const c128*>(in);
  c128* o = reinterpret_cast<c128*>(out);
  if (mode == 0) { for (int64_t n = 0; n < ncx * ncy; ++n) o[n] = restrict_point(n, nx, ny, ncx, a, scale); }
  else { for (int64_t n = 0; n < nx * ny; ++n) o[n] = prolong_point(n, nx, ncx, ncy, a); }
  return FDFD_OK;
}

// the same two cores run by their device kernels (GPU needed): lets a test compare device and host results entry by entry
extern "C" int fdfd_debug_ml_lsq_gpu(fdfd_ctx* ctx, int k, const fdfd_c128* H, double beta, fdfd_c128* y, double* resnorm) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  ARG_CHECK(ctx, k >= 1 && k <= kMaxK && H && y, "bad arguments");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  DevBuf<c128> dH, dy, dsc, dscr;
  CUDA_TRY(ctx, dH.alloc((size_t)(k + 1) * k)); CUDA_TRY(ctx, dy.alloc(k)); CUDA_TRY(ctx, dsc.alloc(2));
  CUDA_TRY(ctx, dscr.alloc((size_t)k * (k + 1) / 2 + 3 * (size_t)k + 2));
  const c128 hb[2] = {c128(beta, 0.0), c128(0.0, 0.0)};
  FDFD_TRY(fdfd_copy_in(ctx, dH.p, H, sizeof(c128) * (size_t)(k + 1) * k));
  CUDA_TRY(ctx, cudaMemcpyAsync(dsc.p, hb, sizeof(hb), cudaMemcpyHostToDevice, ctx->stream));
  k_ml_lsq<<<1, 32, 0, ctx->stream>>>(k, k + 1, dH.p, dsc.p, dy.p, dsc.p + 1, dscr.p, k); KLAUNCH(ctx);
  CUDA_TRY(ctx, cudaGetLastError());
  c128 hr[2];
  CUDA_TRY(ctx, cudaMemcpyAsync(hr, dsc.p, sizeof(hr), cudaMemcpyDeviceToHost, ctx->stream));
  FDFD_TRY(fdfd_copy_out(ctx, y, dy.p, sizeof(c128) * k));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (resnorm) *resnorm = hr[1].x;
  return FDFD_OK;
}
extern "C" int fdfd_debug_ml_transfer_gpu(fdfd_ctx* ctx, int64_t nx, int64_t ny, int mode, double scale, const fdfd_c128* in, fdfd_c128* out) {
  ARG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
  ARG_CHECK(ctx, nx >= 2 && ny >= 2 && in && out, "bad arguments");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int64_t ncx = (nx + 1) / 2, ncy = (ny + 1) / 2, Nf = nx * ny, Nc = ncx * ncy;
  DevBuf<c128> din, dout;
  CUDA_TRY(ctx, din.alloc(mode == 0 ? Nf : Nc)); CUDA_TRY(ctx, dout.alloc(mode == 0 ? Nc : Nf));
  FDFD_TRY(fdfd_copy_in(ctx, din.p, in, sizeof(c128) * (size_t)(mode == 0 ? Nf : Nc)));
  if (mode == 0) { k_ml_restrict<<<vec_blocks_for(ctx, Nc), 256, 0, ctx->stream>>>(nx, ny, ncx, ncy, din.p, dout.p, scale); KLAUNCH(ctx); }
  else { k_ml_prolong<<<vec_blocks_for(ctx, Nf), 256, 0, ctx->stream>>>(nx, ny, ncx, ncy, din.p, dout.p); KLAUNCH(ctx); }
  CUDA_TRY(ctx, cudaGetLastError());
  FDFD_TRY(fdfd_copy_out(ctx, out, dout.p, sizeof(c128) * (size_t)(mode == 0 ? Nc : Nf)));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FDFD_OK;
}
