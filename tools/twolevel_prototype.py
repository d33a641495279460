"""Research prototype (NOT product, NOT oracle): the two-level deflated solve with the components the GPU library really
has -- an fp32 multigrid W-cycle for the fine shifted-Laplacian M^-1 (tools/mg_prototype.MG), a FEW steps of BiCGSTAB
preconditioned by the level-1 multigrid cycle for the coarse Helmholtz system E = Z^T A Z, bilinear Z -- inside flexible
GMRES.  Counts what the GPU would pay: fine multigrid cycles, fine operator applies, coarse (level-1) cycles.

    python tools/twolevel_prototype.py N [inner=5] [beta=0.5] [mode=adef1|mult]
"""
import os, sys, time, math
import numpy as np, scipy.sparse as sp
sys.path.insert(0, "/root/repo")
from oracle import fdfd_oracle as O
from tools.mg_prototype import MG, Level, synth_device, bicgstab as bicgstab_mf


def prolong1d(n):
    nc = n // 2
    rows, cols, vals = [], [], []
    for I in range(nc):
        rows += [2 * I, 2 * I + 1, 2 * I + 1]; cols += [I, I, (I + 1) % nc]; vals += [1.0, 0.5, 0.5]
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, nc))


def assemble(L):
    """sparse matrix of a Level operator, C-order flattening of the (Nx,Ny) arrays"""
    n = L.Nx * L.Ny
    idx = np.arange(n).reshape(L.Nx, L.Ny)
    rows, cols, vals = [], [], []
    def add(shift, axis, coef):
        rows.append(idx.ravel()); cols.append(np.roll(idx, shift, axis).ravel()); vals.append(np.broadcast_to(coef, idx.shape).ravel())
    add(1, 0, L.cxm[:, None]); add(-1, 0, L.cxp[:, None]); add(1, 1, L.cym[None, :]); add(-1, 1, L.cyp[None, :])
    rows.append(idx.ravel()); cols.append(idx.ravel()); vals.append(L.diag.ravel())
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


def fgmres(A, b, prec, tol, maxit, m=80, log=True):
    n = b.size; x = np.zeros_like(b); nb = np.linalg.norm(b); total = 0; hist = []
    while total < maxit:
        r = b - A(x); beta = np.linalg.norm(r)
        if beta <= tol * nb: break
        V = [r / beta]; Zs = []; H = np.zeros((m + 1, m), complex); g = np.zeros(m + 1, complex); g[0] = beta
        for k in range(m):
            z = prec(V[k]); Zs.append(z); w = A(z); total += 1
            for i in range(k + 1):
                H[i, k] = np.vdot(V[i], w); w = w - H[i, k] * V[i]
            H[k + 1, k] = np.linalg.norm(w); V.append(w / H[k + 1, k])
            y, *_ = np.linalg.lstsq(H[:k + 2, :k + 1], g[:k + 2], rcond=None)
            res = np.linalg.norm(H[:k + 2, :k + 1] @ y - g[:k + 2]) / nb
            hist.append(res)
            if log and total % 5 == 0: print(f"    outer {total:3d} relres {res:.2e}", flush=True)
            if res <= tol or total >= maxit: break
        for i in range(len(y)): x = x + y[i] * Zs[i]
    return x, total, hist


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    inner = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    beta = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    mode = sys.argv[4] if len(sys.argv) > 4 else "adef1"
    d = synth_device(n, n)
    if os.environ.get("BENCH_MAP"):   # the bench workload's map (density 1/160) instead of the prototype's denser one (1/40)
        import fdfd_jl_b200 as fdfd
        from fdfd_jl_b200 import workloads as wl
        d.eps_r[:] = wl.synthetic_tm_device(fdfd, n, n, density=1 / 160).eps_r
    g = d.grid; omega = d.omega[0]
    eps0, mu0, _ = O.normalize_parameters(g)
    cxm, cxp, cym, cyp = O.stencil_coefficients(g, omega, "fb")
    Aop = Level(n, n, cxm, cxp, cym, cyp, omega ** 2 * eps0 * d.eps_r)
    b = 1j * omega * d.src
    nb = np.linalg.norm(b)
    t0 = time.time()
    mg = MG(g, omega, d.eps_r, beta=beta, wj=0.7, wl=0.6, nu1=1, nu2=1, cycle="W", wdepth=2, coarse_sweeps=2, min_n=8,
            dtype=np.complex128)
    cnt = {"fineM": 0, "fineA": 0, "coarseM": 0, "coarseE": 0}
    def A(x): cnt["fineA"] += 1; return Aop.apply(x)
    def Minv(r): cnt["fineM"] += 1; return mg(r)
    # coarse space: bilinear prolongation, Galerkin coarse Helmholtz operator (9-point), assembled here for brevity
    P1 = prolong1d(n); Z = sp.kron(P1, P1, format="csr")      # C-order (Nx,Ny) flattening: index = ix*Ny + iy -> kron(Px, Py)
    Amat = assemble(Aop)
    E = (Z.T @ Amat @ Z).tocsr()
    nc = n // 2
    def Eop(y): cnt["coarseE"] += 1; return (E @ y.ravel()).reshape(nc, nc)
    if "redisc" in mode:   # NON-Galerkin coarse operator: 4 x the rediscretised (unshifted) level-1 Helmholtz operator
        L1 = mg.levels[1]
        mass1 = (L1.mass / (1 - 1j * beta))
        A1 = Level(nc, nc, L1.cxm, L1.cxp, L1.cym, L1.cyp, mass1)
        def Eop(y): cnt["coarseE"] += 1; return 4.0 * A1.apply(y)
    # the Galerkin operator is Z^T A Z ~ 4 x (rediscretised operator): scale the level-1 cycle accordingly
    def Mc_inv(r): cnt["coarseM"] += 1; return mg.cyc(1, r / 4.0)
    print(f"n={n} inner={inner} beta={beta} mode={mode}: setup {time.time() - t0:.1f}s", flush=True)

    def coarse_solve(gc):
        if inner == 0:   # one cycle, no Krylov
            return Mc_inv(gc)
        y, it, rn = bicgstab_mf(Eop, gc, Mc_inv, tol=1e-3, maxit=inner)
        return y
    def Q(v): return (Z @ coarse_solve((Z.T @ v.ravel()).reshape(nc, nc)).ravel()).reshape(n, n)
    if mode.startswith("adef1"):     # T = M^-1 (I - A Q) + Q
        def T(v):
            q = Q(v)
            return Minv(v - A(q)) + q
    else:                   # multiplicative, smoother first: y = M^-1 v ; y += Q (v - A y)
        def T(v):
            y = Minv(v)
            return y + Q(v - A(y))
    if mode == "csl":
        t = time.time()
        x, it, rn = bicgstab_mf(A, b, Minv, tol=1e-10, maxit=4000)
        print(f"CSL-only BiCGSTAB: {it} iterations, counts {cnt}, true relres {np.linalg.norm(b - Aop.apply(x)) / nb:.2e} ({time.time() - t:.0f}s)")
        return
    t = time.time()
    x, k, hist = fgmres(A, b, T, 1e-10, 300)
    print(f"two-level FGMRES: outer {k}, counts {cnt}, true relres {np.linalg.norm(b - Aop.apply(x)) / nb:.2e} ({time.time() - t:.0f}s)")
    # cost model in units of one fine multigrid cycle (~1 fine M^-1): a level-1 cycle ~ 0.3 (bandwidth 1/4 + latency), A ~ 0.15
    cost = cnt["fineM"] + 0.15 * cnt["fineA"] + 0.3 * cnt["coarseM"] + 0.05 * cnt["coarseE"]
    print(f"cost ~ {cost:.0f} fine-cycle equivalents")


if __name__ == "__main__":
    main()
