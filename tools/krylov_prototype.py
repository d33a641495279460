"""Research prototype (NOT product, NOT oracle): preconditioner applications needed by BiCGSTAB / IDR(s) / GMRES on the
synthetic TM map with the shifted-Laplacian preconditioner inverted exactly (SuperLU).  Decides the Krylov method of K7."""
import sys, time, math
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, "/root/repo")
from oracle import fdfd_oracle as O
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import workloads as wl

def problem(n, density=1/160, beta=0.5):
    d = wl.synthetic_tm_device(fdfd, n, n, density=density)
    go = O.Grid2D(0.02, [15, 15], [0.0, n * 0.02], [0.0, n * 0.02])
    do = O.Device(go, list(d.omega)); do.eps_r[:] = d.eps_r; do.src[:] = d.src
    w = d.omega[0]
    A, b = O.system_matrix(do, w, O.TM)[:2]
    eps0 = O.EPS0 * go.L0
    M = A - 1j * beta * w * w * eps0 * sp.diags(d.eps_r.ravel(order="F"))
    return A.tocsr(), np.asarray(b).ravel(), spla.splu(M.tocsc())

def bicgstab(A, b, Minv, tol, maxit):
    x = np.zeros_like(b); r = b.copy(); rh = r.copy(); rho = alpha = om = 1.0; v = p = np.zeros_like(b); nb = np.linalg.norm(b); napp = 0
    for it in range(maxit):
        rho1 = np.vdot(rh, r); beta = (rho1 / rho) * (alpha / om); rho = rho1
        p = r + beta * (p - om * v); ph = Minv(p); v = A @ ph; napp += 1
        alpha = rho / np.vdot(rh, v); s = r - alpha * v
        sh = Minv(s); t = A @ sh; napp += 1
        om = np.vdot(t, s) / np.vdot(t, t); x += alpha * ph + om * sh; r = s - om * t
        if np.linalg.norm(r) <= tol * nb: break
    return x, napp

def idrs(A, b, Minv, s, tol, maxit, seed=0):
    """IDR(s) biortho variant (van Gijzen & Sonneveld, ACM TOMS 2011), right-preconditioned."""
    n = len(b); rng = np.random.default_rng(seed)
    P = rng.standard_normal((n, s)) + 1j * rng.standard_normal((n, s)); P, _ = np.linalg.qr(P)
    x = np.zeros_like(b); r = b.copy(); nb = np.linalg.norm(b); napp = 0
    G = np.zeros((n, s), complex); U = np.zeros((n, s), complex); Mm = np.eye(s, dtype=complex); om = 1.0
    while napp < maxit:
        f = P.conj().T @ r
        for k in range(s):
            c = np.linalg.solve(Mm[k:, k:], f[k:])
            v = r - G[:, k:] @ c
            vh = Minv(v)
            U[:, k] = U[:, k:] @ c + om * vh
            G[:, k] = A @ U[:, k]; napp += 1
            for i in range(k):
                a = np.vdot(P[:, i], G[:, k]) / Mm[i, i]
                G[:, k] -= a * G[:, i]; U[:, k] -= a * U[:, i]
            Mm[k:, k] = P[:, k:].conj().T @ G[:, k]
            bta = f[k] / Mm[k, k]
            r = r - bta * G[:, k]; x = x + bta * U[:, k]
            if np.linalg.norm(r) <= tol * nb: return x, napp
            if k + 1 < s: f[k + 1:] = f[k + 1:] - bta * Mm[k + 1:, k]
        vh = Minv(r); t = A @ vh; napp += 1
        om = np.vdot(t, r) / np.vdot(t, t)
        # "maintaining the convergence" omega safeguard
        rho = abs(np.vdot(t, r)) / (np.linalg.norm(t) * np.linalg.norm(r))
        if rho < 0.7: om = om * 0.7 / rho
        x = x + om * vh; r = r - om * t
        if np.linalg.norm(r) <= tol * nb: break
    return x, napp

def gmres_count(A, b, Minv, tol, restart, maxit):
    cnt = [0]
    def mv(v): cnt[0] += 1; return A @ Minv(v)
    n = len(b)
    y, info = spla.gmres(spla.LinearOperator((n, n), matvec=mv, dtype=complex), b, rtol=tol, restart=restart, maxiter=maxit)
    return Minv(y), cnt[0], info

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    t0 = time.time(); A, b, lu = problem(n); print(f"n={n} setup {time.time()-t0:.1f}s", flush=True)
    Minv = lu.solve
    tol = 1e-10
    def report(name, x, napp, t):
        print(f"{name:12s} precond applications={napp:5d}  true relres={np.linalg.norm(b - A @ x)/np.linalg.norm(b):.2e}  ({t:.0f}s)", flush=True)
    t = time.time(); x, k = bicgstab(A, b, Minv, tol, 5000); report("BiCGSTAB", x, k, time.time() - t)
    for s in (2, 4, 8):
        t = time.time(); x, k = idrs(A, b, Minv, s, tol, 10000); report(f"IDR({s})", x, k, time.time() - t)
    for m in (20, 50, 400):
        t = time.time(); x, k, info = gmres_count(A, b, Minv, tol, m, 40); report(f"GMRES({m})", x, k, time.time() - t)
