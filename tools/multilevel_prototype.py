"""Research prototype (NOT product, NOT oracle): multilevel Krylov (Erlangga & Nabben 2008; Sheikh et al. 2016) with the
library's real components: on every level l a flexible GMRES on the REDISCRETISED Helmholtz operator A_l (the multigrid
hierarchy's level operator with the complex shift removed), preconditioned by
    T_l v = q + M_l^-1 (v - A_l q),   q = Z_l (A_{l+1})^-1 Z_l^T v / 4      (ADEF-1, non-Galerkin coarse operator)
where M_l^-1 is ONE multigrid cycle started on level l and (A_{l+1})^-1 is spec[l+1] iterations of the same method one
level down; the last level in `spec` runs BiCGSTAB + multigrid cycle only.

    python tools/multilevel_prototype.py N outer,k1[,k2...]      e.g.  1024 300,8,6
Prints the number of multigrid cycles started on every level (what the GPU pays) and a bandwidth cost estimate.
"""
import os, sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, "/root/repo")
from oracle import fdfd_oracle as O
from tools.mg_prototype import MG, Level, synth_device, bicgstab as bicgstab_mf


T0 = time.time()


def prolong1d(n):
    nc = n // 2
    rows, cols, vals = [], [], []
    for I in range(nc):
        rows += [2 * I, 2 * I + 1, 2 * I + 1]; cols += [I, I, (I + 1) % nc]; vals += [1.0, 0.5, 0.5]
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, nc))


def prolong1d_cubic(n):
    """even fine points inject, odd ones take the 4-point cubic stencil (-1, 9, 9, -1)/16 (periodic)"""
    nc = n // 2
    rows, cols, vals = [], [], []
    for I in range(nc):
        rows += [2 * I]; cols += [I]; vals += [1.0]
        for off, w in ((-1, -1 / 16), (0, 9 / 16), (1, 9 / 16), (2, -1 / 16)):
            rows.append(2 * I + 1); cols.append((I + off) % nc); vals.append(w)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, nc))


def fgmres_fixed(A, b, prec, tol, maxit, restart=60, log=False):
    """flexible GMRES, stops at tol (relative to ||b||) or after maxit preconditioner applications"""
    x = np.zeros_like(b); nb = np.linalg.norm(b); total = 0
    if nb == 0: return x, 0
    while total < maxit:
        r = b - A(x) if total else b.copy(); beta = np.linalg.norm(r)
        if beta <= tol * nb: break
        m = min(restart, maxit - total)
        V = [r / beta]; Zs = []; H = np.zeros((m + 1, m), complex); g = np.zeros(m + 1, complex); g[0] = beta
        for k in range(m):
            z = prec(V[k]); Zs.append(z); w = A(z); total += 1
            if os.environ.get("CGS") and not log:   # classical Gram-Schmidt, one pass (inner levels only): all dots against the same w
                h = [np.vdot(V[i], w) for i in range(k + 1)]
                for i in range(k + 1): H[i, k] = h[i]; w = w - h[i] * V[i]
            else:
                for i in range(k + 1):
                    H[i, k] = np.vdot(V[i], w); w = w - H[i, k] * V[i]
            H[k + 1, k] = np.linalg.norm(w); V.append(w / H[k + 1, k])
            y, *_ = np.linalg.lstsq(H[:k + 2, :k + 1], g[:k + 2], rcond=None)
            res = np.linalg.norm(H[:k + 2, :k + 1] @ y - g[:k + 2]) / nb
            if log: print(f"    outer {total:3d} relres {res:.2e}  ({time.time() - T0:.0f}s)", flush=True)
            if res <= tol: break
        for i in range(len(y)): x = x + y[i] * Zs[i]
        if res <= tol: break
    return x, total


def main():
    n = int(sys.argv[1]); spec = [int(a) for a in sys.argv[2].split(",")]
    beta = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    inner_tol = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
    d = synth_device(n, n)
    if os.environ.get("BENCH_MAP"):   # the bench workload's map (density 1/160) instead of the prototype's denser one (1/40)
        import fdfd_jl_b200 as fdfd
        from fdfd_jl_b200 import workloads as wl
        d.eps_r[:] = wl.synthetic_tm_device(fdfd, n, n, density=1 / 160).eps_r
    g = d.grid; omega = d.omega[0]
    eps0, mu0, _ = O.normalize_parameters(g)
    cxm, cxp, cym, cyp = O.stencil_coefficients(g, omega, "fb")
    b = 1j * omega * d.src
    mg = MG(g, omega, d.eps_r, beta=beta, wj=0.7, wl=0.6, nu1=1, nu2=1, cycle="W", wdepth=2, coarse_sweeps=2, min_n=8, dtype=np.complex128)
    nlev = len(spec)
    ops = [Level(n, n, cxm, cxp, cym, cyp, omega ** 2 * eps0 * d.eps_r)]      # level 0: the reference operator itself
    for l in range(1, nlev):
        L = mg.levels[l]
        ops.append(Level(L.Nx, L.Ny, L.cxm, L.cxp, L.cym, L.cyp, L.mass / (1 - 1j * beta)))
    P1 = prolong1d_cubic if os.environ.get("CUBIC") else prolong1d
    Zs = [sp.kron(P1(ops[l].Nx), P1(ops[l].Ny), format="csr") for l in range(nlev - 1)]
    cntM = [0] * nlev; cntA = [0] * nlev

    def A(l):
        def f(x): cntA[l] += 1; return ops[l].apply(x)
        return f
    def Minv(l):
        def f(r): cntM[l] += 1; return mg.cyc(l, r)
        return f

    def solve(l, rhs, tol):
        if l == nlev - 1:
            if spec[l] == 0: return Minv(l)(rhs)
            if spec[l] < 0:   # bottom level by flexible GMRES(|k|) preconditioned with the cycle only (what the GPU version runs)
                y, _ = fgmres_fixed(A(l), rhs, Minv(l), max(tol, 1e-12), -spec[l]); return y
            y, it, rn = bicgstab_mf(A(l), rhs, Minv(l), tol=max(tol, 1e-12), maxit=spec[l])
            return y
        nx, ny = ops[l].Nx, ops[l].Ny
        def T(v):
            gc = (Zs[l].T @ v.ravel()).reshape(nx // 2, ny // 2) / 4.0
            q = (Zs[l] @ solve(l + 1, gc, inner_tol).ravel()).reshape(nx, ny)
            return q + Minv(l)(v - A(l)(q))
        x, k = fgmres_fixed(A(l), rhs, T, tol, spec[l], log=(l == 0 and bool(os.environ.get("PROGRESS"))))
        if l == 0: print(f"  outer iterations: {k}", flush=True)
        return x

    t = time.time()
    x = solve(0, b, 1e-10)
    rel = np.linalg.norm(b - ops[0].apply(x)) / np.linalg.norm(b)
    # bandwidth cost in fine multigrid cycles: a cycle started on level l moves 4^-l of the data, an apply ~0.15 of a cycle
    cost = sum((cntM[l] + 0.15 * cntA[l]) * 4.0 ** -l for l in range(nlev))
    # latency-aware variant: a cycle on level l never costs less than 0.15 of a fine cycle (coarse levels are launch-bound)
    cost_lat = sum(cntM[l] * max(4.0 ** -l, 0.15) + 0.15 * cntA[l] * 4.0 ** -l for l in range(nlev))
    print(f"n={n} spec={spec} beta={beta} inner_tol={inner_tol}: true relres {rel:.2e}, cycles per level {cntM}, applies per level {cntA}, "
          f"cost ~ {cost:.0f} (bandwidth) / {cost_lat:.0f} (latency floor) fine-cycle equivalents  ({time.time() - t:.0f}s)", flush=True)


if __name__ == "__main__":
    main()
