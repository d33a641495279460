"""Research prototype (NOT product, NOT oracle): how do the two solvers scale with the LENGTH of a straight waveguide?
Config 5's modulated guide stagnates on 8 slabs and BiCGSTAB's iteration count on that device is proportional to the guide length
(profiles/r02_modulated_long_guide.log).  Here: driven TM, eps = 12 guide of width 0.3 um along x, Nx x 128 cells of 0.02 um, x-normal
line source; BiCGSTAB + multigrid (W depth 3) against the multilevel Krylov method (F cycles, steps 6,6), both with the library's real
multigrid components as restated in tools/mg_prototype.py (corner treatment = mean of the y-line and x-line update, like the CUDA code).
    python tools/longguide_prototype.py 512 1024 2048 4096"""
import os, sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fdfd_oracle as O
import tools.mg_prototype as mp
from tools.mg_prototype import MG, Level, bicgstab
from tools.multilevel_prototype import fgmres_fixed, prolong1d


def smooth_mean_corners(self, L, u, f, nn):
    """the CUDA smoother: one residual per sweep; Jacobi outside the strips, y-lines on strip columns, x-lines on strip rows, corners
    take the mean of both line updates"""
    for _ in range(nn):
        r = f - L.apply(u)
        du = np.where(L.ptmask, self.wj * r / L.diag, 0)
        duy = np.zeros_like(u); dux = np.zeros_like(u)
        for k, ix in enumerate(L.xcols): duy[ix, :] = self.wl * L.ylu[k].solve(r[ix, :])
        for k, iy in enumerate(L.yrows): dux[:, iy] = self.wl * L.xlu[k].solve(r[:, iy])
        both = np.zeros(u.shape, bool)
        if len(L.xcols) and len(L.yrows): both[np.ix_(L.xcols, L.yrows)] = True
        d2 = duy + dux; d2[both] *= 0.5
        u = u + du + d2
    return u


MG.smooth = smooth_mean_corners


def run(Nx, Ny=128):
    w = 2 * np.pi * 200e12
    g = O.Grid2D(0.02, [15, 15], [0, Nx * 0.02], [-Ny * 0.01, Ny * 0.01])
    d = O.Device(g, [w])
    O.mask_values(d.eps_r, g, lambda x, y: abs(y) <= 0.15, 12.0)
    d.src[25, np.abs(O.yc(g)) <= 0.6] = 1j
    eps0, mu0, _ = O.normalize_parameters(g)
    cxm, cxp, cym, cyp = O.stencil_coefficients(g, w, "fb")
    A0 = Level(Nx, Ny, cxm, cxp, cym, cyp, w ** 2 * eps0 * d.eps_r)
    b = 1j * w * d.src
    out = {}
    t = time.time()
    mgw = MG(g, w, d.eps_r, beta=0.5, wj=0.7, wl=0.6, nu1=1, nu2=1, cycle="W", wdepth=3, coarse_sweeps=2, min_n=8)
    x, it, rn = bicgstab(A0.apply, b, mgw, maxit=6000)
    out["bicgstab_iters"] = it; out["bicgstab_relres"] = rn; out["bicgstab_s"] = time.time() - t
    t = time.time()
    mgf = MG(g, w, d.eps_r, beta=0.5, wj=0.7, wl=0.6, nu1=1, nu2=1, cycle="F", coarse_sweeps=2, min_n=8)
    spec = [400, 6, 6]
    ops = [A0]
    for l in range(1, 3):
        L = mgf.levels[l]
        ops.append(Level(L.Nx, L.Ny, L.cxm, L.cxp, L.cym, L.cyp, L.mass / (1 - 0.5j)))
    Zs = [sp.kron(prolong1d(ops[l].Nx), prolong1d(ops[l].Ny), format="csr") for l in range(2)]
    cnt = [0, 0, 0]

    def solve(l, rhs, tol):
        Aop = ops[l].apply
        def Minv(r): cnt[l] += 1; return mgf.cyc(l, r)
        if l == 2:
            return fgmres_fixed(Aop, rhs, Minv, 1e-12, spec[l])[0]
        nx, ny = ops[l].Nx, ops[l].Ny
        def T(v):
            gc = (Zs[l].T @ v.ravel()).reshape(nx // 2, ny // 2) / 4.0
            q = (Zs[l] @ solve(l + 1, gc, 0.0).ravel()).reshape(nx, ny)
            return q + Minv(v - Aop(q))
        x, k = fgmres_fixed(Aop, rhs, T, tol, spec[l], restart=96)
        if l == 0: out["ml_outer"] = k
        return x
    x = solve(0, b, 1e-10)
    out["ml_relres"] = np.linalg.norm(b - A0.apply(x)) / np.linalg.norm(b)
    out["ml_cycles"] = list(cnt); out["ml_s"] = time.time() - t
    return out


if __name__ == "__main__":
    for Nx in [int(a) for a in sys.argv[1:]] or [512, 1024]:
        r = run(Nx)
        print(f"guide {Nx} x 128 ({Nx * 0.02 / (1.5 / 3.46):.0f} wavelengths in eps = 12): BiCGSTAB+MG(W3) {r['bicgstab_iters']} iterations "
              f"(relres {r['bicgstab_relres']:.1e}, {r['bicgstab_s']:.0f} s);  multilevel (F, 6,6) {r['ml_outer']} outer iterations, cycles per level {r['ml_cycles']} "
              f"(relres {r['ml_relres']:.1e}, {r['ml_s']:.0f} s)", flush=True)
