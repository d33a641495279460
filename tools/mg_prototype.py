"""Research prototype (NOT product, NOT oracle): calibrates the Krylov + shifted-Laplacian
multigrid design on the CPU before it is written in CUDA.  NumPy, matrix-free, periodic.

Operator form shared by every level (5-point, 1-D PML coefficient arrays + 2-D mass term):
    (A u)[ix,iy] = cxm[ix](u[ix-1]-u) + cxp[ix](u[ix+1]-u) + cym[iy](u[iy-1]-u) + cyp[iy](u[iy+1]-u) + m[ix,iy] u
"""
import sys, time, math
import numpy as np

sys.path.insert(0, "/root/repo")
from oracle import fdfd_oracle as O


def s_profile(p, N, Npml, dw, omega, eps0, m=3.5, lnR=-12.0):
    """continuous s-factor at (1-based, possibly half-integer) position p; pml.jl:1-31 generalised."""
    p = np.asarray(p, dtype=float)
    if Npml == 0:
        return np.ones(p.shape, complex)
    Tw = Npml * dw
    smax = -(m + 1) * lnR / (2 * O.ETA0 * Tw)
    # periodic positions: wrap into [1, N+1)
    p = (p - 1) % N + 1
    depth = np.maximum(np.maximum((Npml + 1) - p, p - (N - Npml + 1)), 0.0) * dw
    return 1 - 1j * smax * (depth / Tw) ** m / (omega * eps0)


class Level:
    def __init__(self, Nx, Ny, cxm, cxp, cym, cyp, mass):
        self.Nx, self.Ny = Nx, Ny
        self.cxm, self.cxp, self.cym, self.cyp = cxm, cxp, cym, cyp
        self.mass = mass
        self.diag = mass - (cxm + cxp)[:, None] - (cym + cyp)[None, :]

    def setup_lines(self, npx, npy):
        """precompute LU of the cyclic tridiagonal systems of the PML strips: y-lines (columns ix in x-PML)
        and x-lines (rows iy in y-PML)."""
        import scipy.sparse as sp, scipy.sparse.linalg as spla
        self.xcols = np.r_[0:npx, self.Nx - npx:self.Nx] if npx > 0 else np.zeros(0, int)
        self.yrows = np.r_[0:npy, self.Ny - npy:self.Ny] if npy > 0 else np.zeros(0, int)
        def cyc(lo, di, up):
            n = len(di); i = np.arange(n)
            return spla.splu(sp.csc_matrix((np.r_[lo, di, up], (np.r_[i, i, i], np.r_[(i - 1) % n, i, (i + 1) % n])), shape=(n, n)))
        self.ylu = [cyc(self.cym, self.diag[ix, :], self.cyp) for ix in self.xcols]
        self.xlu = [cyc(self.cxm, self.diag[:, iy], self.cxp) for iy in self.yrows]
        self.ptmask = np.ones((self.Nx, self.Ny), bool)
        self.ptmask[self.xcols, :] = False; self.ptmask[:, self.yrows] = False

    def apply(self, u):
        return (self.cxm[:, None] * np.roll(u, 1, 0) + self.cxp[:, None] * np.roll(u, -1, 0)
                + self.cym[None, :] * np.roll(u, 1, 1) + self.cyp[None, :] * np.roll(u, -1, 1)
                + self.diag * u)


def coeffs_1d(N, Npml, dw, omega, eps0, mu0, stride, ordering="fb"):
    """rediscretised coefficients at a level whose points sit at fine positions 1 + I*stride."""
    n = N // stride
    I = np.arange(n)
    sb = 1.0 / s_profile(1 + I * stride, N, Npml, dw, omega, eps0)            # at points
    sf = 1.0 / s_profile(1 + I * stride + stride / 2, N, Npml, dw, omega, eps0)  # at midpoints
    h = dw * stride
    if ordering == "fb":
        cm = sf * sb / (mu0 * h * h)
        cp = sf * np.roll(sb, -1) / (mu0 * h * h)
    else:
        cp = sb * sf / (mu0 * h * h)
        cm = sb * np.roll(sf, 1) / (mu0 * h * h)
    return cm, cp


def restrict_fw(r):
    """full weighting, vertex-centred periodic: coarse I <- fine 2I."""
    rx = 0.25 * np.roll(r, 1, 0) + 0.5 * r + 0.25 * np.roll(r, -1, 0)
    rx = rx[::2, :]
    ry = 0.25 * np.roll(rx, 1, 1) + 0.5 * rx + 0.25 * np.roll(rx, -1, 1)
    return ry[:, ::2]


def prolong_bilinear(e, Nx, Ny):
    out = np.zeros((Nx, e.shape[1]), complex)
    out[::2, :] = e
    out[1::2, :] = 0.5 * (e + np.roll(e, -1, 0))
    out2 = np.zeros((Nx, Ny), complex)
    out2[:, ::2] = out
    out2[:, 1::2] = 0.5 * (out + np.roll(out, -1, 1))
    return out2


class MG:
    def __init__(self, g, omega, eps_r, beta=0.5, ordering="fb", min_n=16, max_levels=20,
                 wj=0.8, nu1=1, nu2=1, coarse_sweeps=None, cycle="V", dtype=np.complex128, lines=True, wl=0.7, pad=1, wdepth=99):
        eps0, mu0, _ = O.normalize_parameters(g)
        Nx, Ny = g.size()
        self.levels = []
        mass = (1 - 1j * beta) * omega ** 2 * eps0 * eps_r
        stride = 1
        nx, ny = Nx, Ny
        while True:
            cxm, cxp = coeffs_1d(Nx, g.Npml[0], O.dx(g), omega, eps0, mu0, stride, ordering)
            cym, cyp = coeffs_1d(Ny, g.Npml[1], O.dy(g), omega, eps0, mu0, stride, ordering)
            self.levels.append(Level(nx, ny, cxm.astype(dtype), cxp.astype(dtype), cym.astype(dtype),
                                     cyp.astype(dtype), mass.astype(dtype)))
            if nx % 2 or ny % 2 or nx // 2 < min_n or ny // 2 < min_n or len(self.levels) >= max_levels:
                break
            mass = restrict_fw(mass)
            nx //= 2; ny //= 2; stride *= 2
        self.wj, self.nu1, self.nu2, self.cycle = wj, nu1, nu2, cycle
        self.lines, self.wl = lines, wl
        self.wdepth = wdepth
        if lines:
            for l, L in enumerate(self.levels):
                st = 2 ** l
                L.setup_lines(min(L.Nx // 2, -(-g.Npml[0] // st) + pad) if g.Npml[0] else 0,
                              min(L.Ny // 2, -(-g.Npml[1] // st) + pad) if g.Npml[1] else 0)
        self.coarse_sweeps = coarse_sweeps
        self.dtype = dtype
        self.coarse_lu = None
        if coarse_sweeps is None:
            import scipy.sparse as sp, scipy.sparse.linalg as spla
            L = self.levels[-1]
            n = L.Nx * L.Ny
            idx = np.arange(n).reshape(L.Nx, L.Ny)
            rows, cols, vals = [], [], []
            def add(shift, axis, coef):
                rows.append(idx.ravel()); cols.append(np.roll(idx, shift, axis).ravel()); vals.append(np.broadcast_to(coef, idx.shape).ravel())
            add(1, 0, L.cxm[:, None]); add(-1, 0, L.cxp[:, None]); add(1, 1, L.cym[None, :]); add(-1, 1, L.cyp[None, :])
            rows.append(idx.ravel()); cols.append(idx.ravel()); vals.append(L.diag.ravel())
            A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
            self.coarse_lu = spla.splu(A)
        print("MG levels:", [(l.Nx, l.Ny) for l in self.levels], file=sys.stderr)

    def smooth(self, L, u, f, n):
        if not self.lines:
            for _ in range(n):
                u = u + self.wj * (f - L.apply(u)) / L.diag
            return u
        for _ in range(n):
            r = f - L.apply(u)
            du = np.where(L.ptmask, self.wj * r / L.diag, 0)
            for k, ix in enumerate(L.xcols):
                du[ix, :] = (self.wl * L.ylu[k].solve(r[ix, :].astype(np.complex128))).astype(self.dtype)
            u = u + du
            if len(L.yrows):
                r = f - L.apply(u)
                du = np.zeros_like(u)
                for k, iy in enumerate(L.yrows):
                    du[:, iy] = (self.wl * L.xlu[k].solve(r[:, iy].astype(np.complex128))).astype(self.dtype)
                u = u + du
        return u

    def cyc(self, l, f):
        L = self.levels[l]
        if l == len(self.levels) - 1:
            if self.coarse_lu is not None:
                return self.coarse_lu.solve(f.ravel().astype(np.complex128)).reshape(L.Nx, L.Ny).astype(self.dtype)
            return self.smooth(L, np.zeros_like(f), f, self.coarse_sweeps)
        u = self.smooth(L, np.zeros_like(f), f, self.nu1)
        r = f - L.apply(u)
        rc = restrict_fw(r)
        ec = self.cyc(l + 1, rc)
        if (self.cycle == "W" and l < self.wdepth) or (self.cycle == "F" and l > 0):
            rc2 = rc - self.levels[l + 1].apply(ec)
            ec = ec + (self.cyc(l + 1, rc2) if self.cycle == "W" else self.cycV(l + 1, rc2))
        u = u + prolong_bilinear(ec, L.Nx, L.Ny)
        return self.smooth(L, u, f, self.nu2)

    def cycV(self, l, f):
        c = self.cycle; self.cycle = "V"; out = self.cyc(l, f); self.cycle = c; return out

    def __call__(self, f):
        return self.cyc(0, f.astype(self.dtype)).astype(np.complex128)


def bicgstab(A, b, M, tol=1e-10, maxit=2000, log=None):
    x = np.zeros_like(b); r = b.copy(); rh = r.copy()
    bn = np.linalg.norm(b); rho = alpha = om = 1.0
    v = np.zeros_like(b); p = np.zeros_like(b)
    for it in range(1, maxit + 1):
        rho1 = np.vdot(rh, r)
        beta = (rho1 / rho) * (alpha / om)
        p = r + beta * (p - om * v)
        ph = M(p); v = A(ph)
        alpha = rho1 / np.vdot(rh, v)
        s = r - alpha * v
        if np.linalg.norm(s) / bn < tol:
            x += alpha * ph
            return x, it - 0.5, np.linalg.norm(s) / bn
        sh = M(s); t = A(sh)
        om = np.vdot(t, s) / np.vdot(t, t)
        x += alpha * ph + om * sh
        r = s - om * t
        rho = rho1
        rn = np.linalg.norm(r) / bn
        if log and it % log == 0:
            print(f"  it {it} relres {rn:.3e}", file=sys.stderr)
        if rn < tol:
            return x, it, rn
        if not np.isfinite(rn):
            return x, it, rn
    return x, maxit, rn


def gmres(A, b, M, tol=1e-10, restart=50, maxit=2000, log=None):
    """right-preconditioned restarted GMRES (flexible: stores Z)."""
    x = np.zeros_like(b); bn = np.linalg.norm(b); total = 0
    while total < maxit:
        r = b - A(x); beta = np.linalg.norm(r)
        if beta / bn < tol: return x, total, beta / bn
        V = [r / beta]; Z = []; H = np.zeros((restart + 1, restart), complex)
        g = np.zeros(restart + 1, complex); g[0] = beta
        cs = np.zeros(restart, complex); sn = np.zeros(restart, complex)
        k_done = 0
        for k in range(restart):
            z = M(V[k]); Z.append(z); w = A(z); total += 1
            for i in range(k + 1):
                H[i, k] = np.vdot(V[i], w); w = w - H[i, k] * V[i]
            H[k + 1, k] = np.linalg.norm(w); V.append(w / H[k + 1, k])
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -np.conj(sn[i]) * H[i, k] + cs[i] * H[i + 1, k]; H[i, k] = t
            d = math.hypot(abs(H[k, k]), abs(H[k + 1, k]))
            cs[k] = abs(H[k, k]) / d if H[k, k] != 0 else 0
            sn[k] = (H[k, k] / abs(H[k, k])) * np.conj(H[k + 1, k]) / d if H[k, k] != 0 else 1
            H[k, k] = cs[k] * H[k, k] + sn[k] * H[k + 1, k]; H[k + 1, k] = 0
            g[k + 1] = -np.conj(sn[k]) * g[k]; g[k] = cs[k] * g[k]
            k_done = k + 1
            rn = abs(g[k + 1]) / bn
            if log and total % log == 0: print(f"  it {total} relres {rn:.3e}", file=sys.stderr)
            if rn < tol or total >= maxit: break
        y = np.linalg.solve(np.triu(H[:k_done, :k_done]), g[:k_done])
        for i in range(k_done): x = x + y[i] * Z[i]
        if rn < tol: return x, total, rn
    return x, total, rn


def make_device(kind, n=None):
    w = 2 * math.pi * 200e12
    if kind == "dipole":
        g = O.Grid2D(6.0 / n, [15, 15], [-3, 3], [-3, 3]); d = O.Device(g, [w]); O.setup_src_point(d, (0, 0))
    elif kind == "wg":  # 512x128 variant of the 500x100 test waveguide
        g = O.Grid2D(0.02, [15, 10], [0, 10.24], [-1.28, 1.28]); d = O.Device(g, [w])
        O.mask_values(d.eps_r, g, lambda x, y: abs(y) <= 0.15, 12.0); O.setup_src_line(d, (1.0, 0), O.X)
    elif kind == "synth":
        d = synth_device(n, n)
    return d


def synth_device(Nx, Ny, dh=0.02, seed=0, npml=15):
    """bench workload: vacuum + ε=12 waveguide along x + seeded ε∈[2,12.25] boxes/cylinders; line source."""
    w = 2 * math.pi * 200e12
    g = O.Grid2D(dh, [npml, npml], [0, Nx * dh], [0, Ny * dh]); d = O.Device(g, [w])
    xs, ys = O.xc(g)[:, None], O.yc(g)[None, :]
    eps = np.ones((Nx, Ny))
    rng = np.random.default_rng(seed)
    Lx, Ly = Nx * dh, Ny * dh
    nshape = max(4, int(Lx * Ly / 40))
    for k in range(nshape):
        cx, cy = rng.uniform(0.1 * Lx, 0.9 * Lx), rng.uniform(0.1 * Ly, 0.9 * Ly)
        e = rng.uniform(2, 12.25)
        if k % 2 == 0:
            r = rng.uniform(0.3, 1.5); eps[(xs - cx) ** 2 + (ys - cy) ** 2 <= r * r] = e
        else:
            wx, wy = rng.uniform(0.3, 3), rng.uniform(0.3, 3); eps[(abs(xs - cx) <= wx / 2) & (abs(ys - cy) <= wy / 2)] = e
    eps[:, np.abs(O.yc(g) - Ly / 2) <= 0.15] = 12.0
    d.eps_r = eps.astype(complex)
    d.src[npml + 10, :] = 1j
    return d


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("kind"); ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--solver", default="bicgstab"); ap.add_argument("--beta", type=float, default=0.5)
    ap.add_argument("--wj", type=float, default=0.8); ap.add_argument("--nu", type=int, default=1)
    ap.add_argument("--cycle", default="V"); ap.add_argument("--coarse", type=int, default=None)
    ap.add_argument("--minn", type=int, default=16); ap.add_argument("--maxlev", type=int, default=20)
    ap.add_argument("--f32", action="store_true"); ap.add_argument("--maxit", type=int, default=2000)
    ap.add_argument("--restart", type=int, default=50); ap.add_argument("--nolines", action="store_true"); ap.add_argument("--wl", type=float, default=0.7); ap.add_argument("--pad", type=int, default=1); ap.add_argument("--wdepth", type=int, default=99); ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    d = make_device(a.kind, a.n)
    g = d.grid; omega = d.omega[0]
    eps0, mu0, _ = O.normalize_parameters(g)
    cxm, cxp, cym, cyp = O.stencil_coefficients(g, omega, "fb")
    Aop = Level(*g.size(), cxm, cxp, cym, cyp, omega ** 2 * eps0 * d.eps_r)
    b = 1j * omega * d.src
    mg = MG(g, omega, d.eps_r, beta=a.beta, wj=a.wj, nu1=a.nu, nu2=a.nu, cycle=a.cycle, coarse_sweeps=a.coarse,
            min_n=a.minn, max_levels=a.maxlev, lines=not a.nolines, wl=a.wl, pad=a.pad, wdepth=a.wdepth, dtype=np.complex64 if a.f32 else np.complex128)
    t = time.time()
    if a.solver == "bicgstab":
        x, it, rn = bicgstab(Aop.apply, b, mg, maxit=a.maxit, log=20)
    else:
        x, it, rn = gmres(Aop.apply, b, mg, restart=a.restart, maxit=a.maxit, log=20)
    true = np.linalg.norm(b - Aop.apply(x)) / np.linalg.norm(b)
    print(f"{a.kind} {g.size()} solver={a.solver} beta={a.beta} wj={a.wj} nu={a.nu} cyc={a.cycle} coarse={a.coarse} f32={a.f32}: "
          f"iters={it} relres={rn:.2e} true={true:.2e} time={time.time()-t:.1f}s")
    if a.check:
        A, bb, _ = O.system_matrix(d, omega, O.TM)
        xr = O.dolinearsolve(A, bb).reshape(g.size(), order="F")
        print("rel L2 vs direct:", np.linalg.norm(x - xr) / np.linalg.norm(xr),
              " stencil-vs-matrix:", np.linalg.norm(A @ x.ravel(order='F') - Aop.apply(x).ravel(order='F')) / np.linalg.norm(bb))
