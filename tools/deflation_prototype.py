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

if __name__ == "__main__":
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
