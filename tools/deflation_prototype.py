"""Research prototype (NOT product, NOT oracle): how much would a two-level deflation (ADEF-1, Sheikh/Lahaye/Vuik) on top of
the shifted-Laplacian preconditioner cut the Krylov iteration count on the synthetic TM map?  Exact M^-1 and exact coarse
solves (SuperLU), so the numbers are upper bounds on what a GPU multilevel version could reach."""
import sys, time, math
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, "/root/repo")
from tools.krylov_prototype import problem, bicgstab

def prolong1d(n):
    nc = n // 2
    rows, cols, vals = [], [], []
    for I in range(nc):
        rows += [2 * I, 2 * I + 1, 2 * I + 1]; cols += [I, I, (I + 1) % nc]; vals += [1.0, 0.5, 0.5]
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, nc))

def deflation(A, n, levels=1):
    Z = None
    m = n
    for _ in range(levels):
        P1 = prolong1d(m)
        Zl = sp.kron(P1, P1, format="csr")   # x fastest: index = ix + Nx*iy -> kron(Py, Px)
        Z = Zl if Z is None else (Z @ Zl).tocsr()
        m //= 2
    E = (Z.T @ A @ Z).tocsc()
    return Z, spla.splu(E)

if __name__ == "__main__" and not (len(sys.argv) > 2 and sys.argv[2] == "inexact"):
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    tol = 1e-10
    t0 = time.time(); A, b, lu = problem(n); Minv = lu.solve
    print(f"n={n}: setup {time.time()-t0:.1f}s", flush=True)
    nb = np.linalg.norm(b)
    x, k = bicgstab(A, b, Minv, tol, 5000)
    print(f"CSL only            : BiCGSTAB {k} preconditioner applications, relres {np.linalg.norm(b-A@x)/nb:.1e}", flush=True)
    for levels in (1, 2):
        t0 = time.time(); Z, Elu = deflation(A, n, levels)
        Q = lambda v: Z @ Elu.solve(Z.T @ v)
        # right preconditioner T = M^-1 (I - A Q) + Q  (ADEF-1)
        T = lambda v: (lambda q: Minv(v - A @ q) + q)(Q(v))
        x, k = bicgstab(A, b, T, tol, 5000)
        print(f"+ deflation, coarse {n >> levels}^2 (exact): BiCGSTAB {k} applications (each = 1 coarse solve + 2 A + 1 M^-1), relres {np.linalg.norm(b-A@x)/nb:.1e}  [setup {time.time()-t0:.0f}s]", flush=True)


# ---- inexact two-level method: how accurate must the coarse Helmholtz solve be? ---------------------------------------
def fgmres(A, b, prec, tol, maxit, m=60):
    """flexible GMRES (the preconditioner contains an inner iteration, so it changes from step to step)"""
    n = len(b); x = np.zeros_like(b); nb = np.linalg.norm(b); total = 0
    while total < maxit:
        r = b - A @ x; beta = np.linalg.norm(r)
        if beta <= tol * nb: break
        V = np.zeros((n, m + 1), complex); Zs = np.zeros((n, m), complex); H = np.zeros((m + 1, m), complex)
        V[:, 0] = r / beta; g = np.zeros(m + 1, complex); g[0] = beta
        k_used = 0
        for k in range(m):
            Zs[:, k] = prec(V[:, k]); w = A @ Zs[:, k]; total += 1
            for i in range(k + 1):
                H[i, k] = np.vdot(V[:, i], w); w -= H[i, k] * V[:, i]
            H[k + 1, k] = np.linalg.norm(w); V[:, k + 1] = w / H[k + 1, k]
            k_used = k + 1
            y, res, *_ = np.linalg.lstsq(H[:k + 2, :k + 1], g[:k + 2], rcond=None)
            if np.linalg.norm(H[:k + 2, :k + 1] @ y - g[:k + 2]) <= tol * nb or total >= maxit: break
        x = x + Zs[:, :k_used] @ y
    return x, total


def inexact_study(n, inners=(None, 40, 20, 10, 5)):
    A, b, lu = problem(n); Minv = lu.solve; nb = np.linalg.norm(b)
    P1 = prolong1d(n); Z = sp.kron(P1, P1, format="csr")
    E = (Z.T @ A @ Z).tocsc()
    # coarse shifted operator: Galerkin product of the fine shifted operator (what a multigrid hierarchy already holds)
    import fdfd_jl_b200 as fdfd
    from fdfd_jl_b200 import workloads as wl
    from oracle import fdfd_oracle as O
    d = wl.synthetic_tm_device(fdfd, n, n, density=1 / 160)
    w = d.omega[0]; eps0 = O.EPS0 * 1e-6
    Mf = A - 1j * 0.5 * w * w * eps0 * sp.diags(d.eps_r.ravel(order="F"))
    Mc = spla.splu((Z.T @ Mf @ Z).tocsc())
    Elu = spla.splu(E)
    print(f"n={n}: outer FGMRES iterations with the level-1 coarse system solved by `inner` BiCGSTAB steps (exact coarse CSL inverse)")
    for inner in inners:
        cnt = [0]
        def coarse(g):
            if inner is None: return Elu.solve(g)
            y, k = bicgstab(E, g, Mc.solve, 1e-12, inner); cnt[0] += k
            return y
        Q = lambda v: Z @ coarse(Z.T @ v)
        T = lambda v: (lambda q: Minv(v - A @ q) + q)(Q(v))
        x, k = fgmres(A, b, T, 1e-10, 400)
        print(f"  inner={'exact' if inner is None else inner:>5}: outer {k:4d}, coarse-level preconditioner applications {cnt[0]:5d}, relres {np.linalg.norm(b - A @ x)/nb:.1e}", flush=True)


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "inexact":
    inexact_study(int(sys.argv[1]), tuple(None if a == 'exact' else int(a) for a in sys.argv[3:]) or (None, 40, 20, 10, 5))
