"""CPU oracle for the FDFD.jl assembly + linear-solve hot path (NumPy/SciPy).

TEST INFRASTRUCTURE ONLY.  This module is a literal restatement of the reference
algorithm (fancompute/FDFD.jl, Julia) used as the *checker* for the CUDA path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product (``fdfd.jl_b200``) never does.

Parity status: the reference cannot be executed here (no Julia).  The oracle is
PINNED against the only known-answer numbers the reference holds, the three
photon-number outputs of ``notebooks/Example_simulations.ipynb`` cells 21/23/25
(modulated TM waveguide, exercises PML, b.f operator ordering, mode source,
sideband coupling, forward-difference H recovery and the flux integral) --
see ``tests/test_oracle_golden.py``.  The driven f.b ordering, TE and
``eigenfrequency`` have no reference known-answers: for those rows parity is
"unpinned beyond restatement" (DESIGN.md says the same).

Third-party arithmetic that is not under /root/reference (all unpinned there,
``REQUIRE:1-9`` / ``Project.toml:6-16``): SuiteSparse UMFPACK behind Julia's
``lu(A)\\b`` (``src/solver/solver.jl:35``) -> ``scipy.sparse.linalg.splu`` (SuperLU);
ARPACK behind ``Arpack.eigs`` (``src/solver/eigen.jl:25,86,104``) ->
``scipy.sparse.linalg.eigs`` (same ARPACK, same transformed-``which`` semantics).

Layout convention (Julia column-major): arrays are indexed ``a[ix, iy]`` and the
flattening ``a[:]`` makes x the fast index, i.e. ``n = ix + Nx*iy``.  In NumPy this
is ``a.ravel(order="F")`` for an ``(Nx, Ny)`` array.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as _dcfield

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --- constants: src/types.jl:6-9 (must match bit-for-bit) -------------------------
EPS0 = 8.85418782e-12
MU0 = 1.25663706e-6
C0 = math.sqrt(1 / EPS0 / MU0)
ETA0 = math.sqrt(MU0 / EPS0)
DEFAULT_L0 = 1e-6  # src/grid.jl:5

X, Y = 0, 1  # Direction x̂, ŷ (0-based here; src/types.jl:18)
FORWARD, BACKWARD = 0, 1  # src/types.jl:19
TM, TE = 0, 1  # src/types.jl:20


def _jround(x: float) -> int:
    """Julia ``round`` (RoundNearest, ties to even) == Python ``round``."""
    return int(round(x))


# --- Grid: src/grid.jl:7-59 --------------------------------------------------------
@dataclass
class Grid:
    L: tuple
    L0: float
    N: tuple
    Npml: tuple
    bounds: tuple  # ((x0[,y0]), (x1[,y1]))

    @property
    def ndim(self):
        return len(self.N)

    def size(self, i=None):
        if i is None:
            return tuple(self.N)
        return 1 if i >= self.ndim else self.N[i]

    def __len__(self):
        return int(np.prod(self.N))


def Grid2D(dh, Npml, xrange, yrange, L0=DEFAULT_L0) -> Grid:
    """src/grid.jl:26-33: N = Int.(round.(L/dh))."""
    L = (float(xrange[1] - xrange[0]), float(yrange[1] - yrange[0]))
    N = (_jround(L[0] / dh), _jround(L[1] / dh))
    return Grid(L, L0, N, (int(Npml[0]), int(Npml[1])),
                ((float(xrange[0]), float(yrange[0])), (float(xrange[1]), float(yrange[1]))))


def Grid1D(N, xrange, L0=DEFAULT_L0) -> Grid:
    """src/grid.jl:53-59: 1-D grid from a cell count, Npml = 0."""
    L = (float(xrange[1] - xrange[0]),)
    return Grid(L, L0, (int(N),), (0,), ((float(xrange[0]),), (float(xrange[1]),)))


def dx(g: Grid):  # src/grid.jl:68-70
    return (g.bounds[1][0] - g.bounds[0][0]) / g.N[0]


def dy(g: Grid):  # src/grid.jl:72-74
    return (g.bounds[1][1] - g.bounds[0][1]) / g.N[1]


def dh(g: Grid, w):  # src/grid.jl:63-66
    return dx(g) if w == X else dy(g)


def xc(g: Grid):  # src/grid.jl:76-78  b1 + dx*(0.5:1:N)
    return g.bounds[0][0] + dx(g) * (0.5 + np.arange(g.N[0]))


def yc(g: Grid):  # src/grid.jl:80-82
    return g.bounds[0][1] + dy(g) * (0.5 + np.arange(g.N[1]))


def x2ind(g: Grid, x):
    """src/grid.jl:112-117 -- returns the 1-based Julia index."""
    ind = int(_jround((x - g.bounds[0][0]) / g.L[0] * g.size(0)) + 1)
    return min(max(ind, 1), g.N[0])


def y2ind(g: Grid, y):  # src/grid.jl:120-125
    ind = int(_jround((y - g.bounds[0][1]) / g.L[1] * g.size(1)) + 1)
    return min(max(ind, 1), g.N[1])


def coord2ind(g: Grid, pt):  # src/grid.jl:103-109 (1-based)
    if g.ndim == 1:
        return x2ind(g, pt[0])
    return x2ind(g, pt[0]), y2ind(g, pt[1])


# --- derivative operators: src/grid.jl:128-154 ---------------------------------
def _delta1d(n, d, s):
    """Periodic 1-D difference matrix; values built as (1/d)*(+-1.0) like the reference."""
    c = 1 / d
    if s == FORWARD:  # spdiagm(1 => ones, 0 => -ones, -N+1 => [1])
        rows = np.concatenate([np.arange(n - 1), np.arange(n), [n - 1]])
        cols = np.concatenate([np.arange(1, n), np.arange(n), [0]])
        vals = np.concatenate([np.ones(n - 1), -np.ones(n), [1.0]])
    else:  # spdiagm(-1 => -ones, 0 => ones, N-1 => [-1])
        rows = np.concatenate([np.arange(1, n), np.arange(n), [0]])
        cols = np.concatenate([np.arange(n - 1), np.arange(n), [n - 1]])
        vals = np.concatenate([-np.ones(n - 1), np.ones(n), [-1.0]])
    if n == 1:  # degenerate: spdiagm would stack entries; never used by the reference
        raise ValueError("delta on a 1-cell axis")
    return sp.csr_matrix((c * vals, (rows, cols)), shape=(n, n))


def delta(w, s, g: Grid):
    """δ(w, s, g): x-operators kron(I_Ny, δ1D), y-operators kron(δ1D, I_Nx)."""
    Nx = g.size(0)
    Ny = g.size(1) if g.ndim == 2 else 1
    if w == X:
        return sp.kron(sp.identity(Ny, format="csr"), _delta1d(Nx, dx(g), s), format="csr")
    return sp.kron(_delta1d(Ny, dy(g), s), sp.identity(Nx, format="csr"), format="csr")


def grid_average(a, w):
    """src/grid.jl:157-162: (a + circshift(a, +1 along w))/2 => avg[i] = (a[i]+a[i-1])/2."""
    a = np.asarray(a)
    if a.ndim == 1:
        return (a + np.roll(a, 1)) / 2
    if w == X:
        return (a + np.roll(a, 1, axis=0)) / 2
    if w == Y:
        return (a + np.roll(a, 1, axis=1)) / 2
    return a


# --- PML: src/pml.jl:1-63 ------------------------------------------------------
def normalize_parameters(g: Grid):  # src/device.jl:40
    return EPS0 * g.L0, MU0 * g.L0, C0 / g.L0


def create_sfactor(w, s, g: Grid, omega, m=3.5, lnR=-12.0):
    """src/pml.jl:1-31.  Returns the s-factor (NOT inverted) 1-D array along w."""
    eps0, _, _ = normalize_parameters(g)
    dw = dh(g, w)
    Nw = g.size(w)
    Nw_pml = g.Npml[w]
    out = np.ones(Nw, dtype=np.complex128)
    if Nw_pml == 0:
        return out  # reference would form Tw=0 -> σmax=Inf but never uses it
    Tw = Nw_pml * dw
    sigma_max = -(m + 1) * lnR / (2 * ETA0 * Tw)

    def S(l):
        return 1 - 1j * (sigma_max * (l / Tw) ** m) / (omega * eps0)

    for i in range(1, Nw + 1):  # 1-based like the reference
        if s == FORWARD:
            if i <= Nw_pml:
                out[i - 1] = S(dw * (Nw_pml - i + 0.5))
            elif i > Nw - Nw_pml:
                out[i - 1] = S(dw * (i - (Nw - Nw_pml) - 0.5))
        else:
            if i <= Nw_pml:
                out[i - 1] = S(dw * (Nw_pml - i + 1))
            elif i > Nw - Nw_pml:
                out[i - 1] = S(dw * (i - (Nw - Nw_pml) - 1))
    return out


def inv_sfactors(g: Grid, omega):
    """The four 1-D *inverse* s-factor arrays (sxf, sxb, syf, syb): src/pml.jl:46-54 (`.^-1`)."""
    return (1.0 / create_sfactor(X, FORWARD, g, omega),
            1.0 / create_sfactor(X, BACKWARD, g, omega),
            1.0 / create_sfactor(Y, FORWARD, g, omega),
            1.0 / create_sfactor(Y, BACKWARD, g, omega))


def S_create(g: Grid, omega):
    """src/pml.jl:33-63: four N x N sparse diagonals of inverse s-factors."""
    sxf, sxb, syf, syb = inv_sfactors(g, omega)
    Nx, Ny = g.size()
    Sxf = sp.diags(np.tile(sxf, Ny), format="csr")  # x-factors tiled along y
    Sxb = sp.diags(np.tile(sxb, Ny), format="csr")
    Syf = sp.diags(np.repeat(syf, Nx), format="csr")  # y-factors repeated along x
    Syb = sp.diags(np.repeat(syb, Nx), format="csr")
    return Sxf, Sxb, Syf, Syb


def scaled_derivatives(g: Grid, omega):
    """src/driven.jl:28-31 / eigen.jl:75-78: (δxb, δxf, δyb, δyf) with PML row scaling."""
    Sxf, Sxb, Syf, Syb = S_create(g, omega)
    return (Sxb @ delta(X, BACKWARD, g), Sxf @ delta(X, FORWARD, g),
            Syb @ delta(Y, BACKWARD, g), Syf @ delta(Y, FORWARD, g))


# --- Device: src/device.jl -----------------------------------------------------
@dataclass
class Mode:  # src/device.jl:5-11
    pol: int
    dir: int
    neff: float
    pt: tuple
    width: float


@dataclass
class Device:  # src/device.jl:19-35
    grid: Grid
    omega: list
    eps_r: np.ndarray = None
    src: np.ndarray = None
    modes: list = _dcfield(default_factory=list)

    def __post_init__(self):
        if np.isscalar(self.omega):
            self.omega = [float(self.omega)]
        shape = self.grid.size()
        if self.eps_r is None:
            self.eps_r = np.ones(shape, dtype=np.complex128)
        if self.src is None:
            self.src = np.zeros(shape, dtype=np.complex128)


@dataclass
class ModulatedDevice(Device):  # src/solver/modulation.jl:4-26
    Omega: float = 0.0
    nsidebands: int = 0
    sharedpml: bool = True
    deps_r: np.ndarray = None

    def __post_init__(self):
        super().__post_init__()
        if self.deps_r is None:
            self.deps_r = np.zeros(self.grid.size(), dtype=np.complex128)


def mask_values(pixels, g: Grid, region, value):
    """src/device.jl:63-83 (`_mask_values!`): region(x,y)->bool, value scalar or f(x,y)."""
    if g.ndim == 2:
        XX, YY = np.meshgrid(xc(g), yc(g), indexing="ij")
        mask = np.vectorize(region)(XX, YY).astype(bool)
        if callable(value):
            pixels[mask] = np.vectorize(value)(XX, YY)[mask]
        else:
            pixels[mask] = value
    else:
        xs = xc(g)
        mask = np.vectorize(region)(xs).astype(bool)
        pixels[mask] = np.vectorize(value)(xs)[mask] if callable(value) else value


def box_region(center, size):
    """GeometryPrimitives Box([cx,cy,..],[wx,wy,..]) membership (closed) in the xy-plane."""
    cx, cy, wx, wy = center[0], center[1], size[0], size[1]
    return lambda x, y: (abs(x - cx) <= wx / 2) and (abs(y - cy) <= wy / 2)


def cylinder_region(center, radius):
    """GeometryPrimitives Cylinder along z, infinite height."""
    cx, cy = center[0], center[1]
    return lambda x, y: (x - cx) ** 2 + (y - cy) ** 2 <= radius ** 2


def compose_shapes(pixels, g: Grid, shapes):
    """src/device.jl:47-61: first shape in list order that contains the pixel wins."""
    XX, YY = np.meshgrid(xc(g), yc(g), indexing="ij")
    done = np.zeros(pixels.shape, dtype=bool)
    for region, value in shapes:
        mask = np.vectorize(region)(XX, YY).astype(bool) & ~done
        pixels[mask] = np.vectorize(value)(XX, YY)[mask] if callable(value) else value
        done |= mask


def setup_src_point(d: Device, pt):  # src/device.jl:95-98
    ix, iy = coord2ind(d.grid, pt)
    d.src[ix - 1, iy - 1] = 1j


def setup_src_line(d: Device, pt, srcnormal):  # src/device.jl:101-108
    ix, iy = coord2ind(d.grid, pt)
    if srcnormal == X:
        d.src[ix - 1, :] = 1j
    else:
        d.src[:, iy - 1] = 1j


def eigenmode_1d(g1: Grid, eps_r, omega, pol, neff, nev=1):
    """src/solver/eigen.jl:6-29.  <=~100 unknowns, so a dense eigendecomposition picks the
    eigenvalues nearest sigma (what ARPACK shift-invert :LM returns)."""
    eps0, mu0, c0 = normalize_parameters(g1)
    Teps = sp.diags(eps0 * eps_r)
    dxb = delta(X, BACKWARD, g1)
    dxf = delta(X, FORWARD, g1)
    if pol == TM:
        A = omega ** 2 * mu0 * Teps + dxf @ dxb
    else:
        Tepsxinv = sp.diags(1.0 / (eps0 * grid_average(eps_r, X)))
        A = omega ** 2 * mu0 * Teps + Teps @ dxf @ Tepsxinv @ dxb
    sigma = (omega / c0 * neff) ** 2
    w, v = np.linalg.eig(A.toarray())
    order = np.argsort(np.abs(w - sigma))[:nev]
    return np.sqrt(w[order].astype(np.complex128)), v[:, order]


def get_modes(d: Device, pol, omega, neff, nmodes, pt, slicenormal, slicewidth):
    """src/device.jl:124-149.  Returns (beta, vectors, ix (0-based index or array), iy)."""
    g = d.grid
    ix, iy = coord2ind(g, pt)
    if slicenormal == X:
        srcpoints = _jround(slicewidth / dy(g))
    else:
        srcpoints = _jround(slicewidth / dx(g))
    if srcpoints % 2 == 0:
        srcpoints += 1
    M = (srcpoints - 1) // 2
    srcpoints = 2 * M + 1
    if slicenormal == X:
        iy = iy + np.arange(-M, M + 1)
        h = dy(g)
        eps_slice = d.eps_r[ix - 1, iy - 1]
    else:
        ix = ix + np.arange(-M, M + 1)
        h = dx(g)
        eps_slice = d.eps_r[ix - 1, iy - 1]
    g1 = Grid1D(srcpoints, (0.0, srcpoints * h), L0=g.L0)
    beta, vec = eigenmode_1d(g1, eps_slice, omega, pol, neff, nmodes)
    return beta, vec, ix - 1, iy - 1


def setup_mode(d: Device, pol, omega, neff, pt, srcnormal, srcwidth):
    """src/device.jl:118-121: src[slice] += normalize(abs.(vector))."""
    _, vec, ix, iy = get_modes(d, pol, omega, neff, 1, pt, srcnormal, srcwidth)
    v = np.abs(vec[:, 0])
    d.src[ix, iy] += v / np.linalg.norm(v)


def _apply_modes(d: Device, omega):
    """src/solver/driven.jl:15-19: reset src only when modes are used; always TM (TODO in ref)."""
    if len(d.modes) > 0:
        d.src = np.zeros(d.grid.size(), dtype=np.complex128)
    for mode in d.modes:
        setup_mode(d, TM, omega, mode.neff, mode.pt, mode.dir, mode.width)


# --- linear solve seam: src/solver/solver.jl:4-41 ------------------------------
def dolinearsolve(A, b):
    """`lu(A)\\b` (UMFPACK) -> SuperLU."""
    lu = spla.splu(sp.csc_matrix(A))
    return lu.solve(np.asarray(b).ravel())


def _F(a):  # Julia a[:]
    return np.asarray(a).ravel(order="F")


def _pack(g: Grid, *vecs):
    """data.jl:60-63 / 75-78: cat(reshape(.,(Nx,Ny))..., dims=3) -> (Nx,Ny,3)."""
    Nx, Ny = g.size()
    return np.stack([np.reshape(v, (Nx, Ny), order="F") for v in vecs], axis=2)


# --- driven solve: src/solver/driven.jl:4-59 -----------------------------------
def system_matrix(d: Device, omega, pol):
    """Returns (A, b, aux) exactly as driven.jl:21-36 / 45-46 builds them."""
    g = d.grid
    eps0, mu0, _ = normalize_parameters(g)
    Teps = sp.diags(_F(eps0 * d.eps_r), format="csr")
    Tepsxi = sp.diags(1.0 / _F(grid_average(eps0 * d.eps_r, X)), format="csr")
    Tepsyi = sp.diags(1.0 / _F(grid_average(eps0 * d.eps_r, Y)), format="csr")
    dxb, dxf, dyb, dyf = scaled_derivatives(g, omega)
    if pol == TM:
        A = dxf * (1 / mu0) @ dxb + dyf * (1 / mu0) @ dyb + omega ** 2 * Teps
    else:
        # `speye` no longer exists on Julia>=1.0; intent is ω²μ₀·I
        A = dxf @ Tepsxi @ dxb + dyf @ Tepsyi @ dyb + omega ** 2 * mu0 * sp.identity(len(g), format="csr")
    b = 1j * omega * _F(d.src)
    return sp.csr_matrix(A), b, (dxb, dxf, dyb, dyf, Tepsxi, Tepsyi)


def solve(d: Device, pol=TM, linsolve=dolinearsolve):
    """solve(d::Device, pol) -> list of dict(omega, data (Nx,Ny,3)); single dict if one ω."""
    g = d.grid
    _, mu0, _ = normalize_parameters(g)
    out = []
    for omega in d.omega:
        _apply_modes(d, omega)
        A, b, (dxb, dxf, dyb, dyf, Tepsxi, Tepsyi) = system_matrix(d, omega, pol)
        u = linsolve(A, b)
        if pol == TM:
            hx = -1 / 1j / omega / mu0 * (dyb @ u)
            hy = 1 / 1j / omega / mu0 * (dxb @ u)
            out.append({"omega": complex(omega), "data": _pack(g, u, hx, hy), "pol": TM})
        else:
            ex = 1 / 1j / omega * (Tepsyi @ (dyb @ u))
            ey = 1 / 1j / omega * (Tepsxi @ (-(dxb @ u)))
            out.append({"omega": complex(omega), "data": _pack(g, u, ex, ey), "pol": TE})
    return out[0] if len(out) == 1 else out


# --- eigenfrequency: src/solver/eigen.jl:69-115 ---------------------------------
def eigen_matrix(d: Device, pol):
    g = d.grid
    eps0, mu0, _ = normalize_parameters(g)
    omega0 = d.omega[0]
    dxb, dxf, dyb, dyf = scaled_derivatives(g, omega0)
    if pol == TM:
        Ti = sp.diags(1.0 / _F(d.eps_r), format="csr")
        A = Ti @ dxf @ dxb + Ti @ dyf @ dyb
        sigma = -omega0 ** 2 * mu0 * eps0
        aux = (dxb, dxf, dyb, dyf, None, None)
    else:
        Txi = sp.diags(1.0 / _F(grid_average(eps0 * d.eps_r, X)), format="csr")
        Tyi = sp.diags(1.0 / _F(grid_average(eps0 * d.eps_r, Y)), format="csr")
        A = dxf @ Txi @ dxb + dyf @ Tyi @ dyb
        sigma = -omega0 ** 2 * mu0
        aux = (dxb, dxf, dyb, dyf, Txi, Tyi)
    return sp.csr_matrix(A), sigma, aux


def eigen_fields(d: Device, pol, lam, vecs, aux):
    """Post-processing of eigen.jl:87-95 / 105-113 (ω from λ, forward-difference H for TM,
    and the literally-swapped ε averaging for TE E-fields)."""
    g = d.grid
    eps0, mu0, _ = normalize_parameters(g)
    dxb, dxf, dyb, dyf, Txi, Tyi = aux
    fields = []
    if pol == TM:
        om = np.sqrt(-lam.astype(np.complex128) / mu0 / eps0)
        for i in range(len(lam)):
            ez = vecs[:, i]
            hx = -1 / 1j / om[i] / mu0 * (dyf @ ez)
            hy = 1 / 1j / om[i] / mu0 * (dxf @ ez)
            fields.append({"omega": om[i], "data": _pack(g, ez, hx, hy), "pol": TM})
    else:
        om = np.sqrt(-lam.astype(np.complex128) / mu0)
        for i in range(len(lam)):
            hz = vecs[:, i]
            ex = 1 / 1j / om[i] * (Txi @ (dyb @ hz))
            ey = 1 / 1j / om[i] * (Tyi @ (-(dxb @ hz)))
            fields.append({"omega": om[i], "data": _pack(g, hz, ex, ey), "pol": TE})
    return om, fields


def eigenfrequency(d: Device, pol, nev, which="LM", v0=None, ncv=None):
    A, sigma, aux = eigen_matrix(d, pol)
    lam, vecs = spla.eigs(sp.csc_matrix(A), k=nev, sigma=sigma, which=which, v0=v0, ncv=ncv)
    return eigen_fields(d, pol, lam, vecs, aux)


# --- modulated MF-FDFD: src/solver/modulation.jl:35-119 -------------------------
def modulated_system(d: ModulatedDevice, omega):
    g = d.grid
    eps0, mu0, _ = normalize_parameters(g)
    ns = d.nsidebands
    nf = 2 * ns + 1
    N = len(g)
    omegan = omega + d.Omega * np.arange(-ns, ns + 1)
    Teps = sp.diags(_F(eps0 * d.eps_r), format="csr")
    Tdeps = sp.diags(_F(eps0 * d.deps_r), format="csr")
    dxb, dxf = delta(X, BACKWARD, g), delta(X, FORWARD, g)
    dyb, dyf = delta(Y, BACKWARD, g), delta(Y, FORWARD, g)
    b = np.zeros(N * nf, dtype=np.complex128)
    b[ns * N:(ns + 1) * N] = 1j * omega * _F(d.src)
    As, S = [], []
    if d.sharedpml:
        Sxf, Sxb, Syf, Syb = S_create(g, omega)
        S.append((Sxf, Sxb, Syf, Syb))
        A1 = (Sxb @ dxb) / mu0 @ Sxf @ dxf + (Syb @ dyb) / mu0 @ Syf @ dyf  # b.f ordering (:82)
        for j in range(nf):
            As.append(A1 + omegan[j] ** 2 * Teps)
    else:
        for j in range(nf):
            Sxf, Sxb, Syf, Syb = S_create(g, omegan[j])
            S.append((Sxf, Sxb, Syf, Syb))
            As.append((Sxb @ dxb) / mu0 @ Sxf @ dxf + (Syb @ dyb) / mu0 @ Syf @ dyf
                      + omegan[j] ** 2 * Teps)
    if ns > 0:
        Cp = sp.kron(sp.diags(0.5 * omegan[:-1] ** 2, 1), Tdeps.conj(), format="csr")
        Cm = sp.kron(sp.diags(0.5 * omegan[1:] ** 2, -1), Tdeps, format="csr")
        A = sp.block_diag(As, format="csr") + Cp + Cm
    else:
        A = As[0]
    A = sp.csr_matrix(A)
    A.eliminate_zeros()  # Julia sparse `+` drops nothing structurally, kron of zeros stores none
    return A, b, omegan, S, (dxf, dyf)


def solve_modulated(d: ModulatedDevice, linsolve=dolinearsolve):
    """Returns fields[iω][j] (j over sidebands -ns..ns), each dict(omega, data)."""
    g = d.grid
    _, mu0, _ = normalize_parameters(g)
    N = len(g)
    nf = 2 * d.nsidebands + 1
    out = []
    for omega in d.omega:
        _apply_modes(d, omega)
        A, b, omegan, S, (dxf, dyf) = modulated_system(d, omega)
        ez = linsolve(A, b)
        row = []
        for j in range(nf):
            Sxf, _, Syf, _ = S[0 if d.sharedpml else j]
            ezi = ez[j * N:(j + 1) * N]
            hx = -1 / 1j / omegan[j] / mu0 * (Syf @ (dyf @ ezi))  # forward diffs (:112-113)
            hy = 1 / 1j / omegan[j] / mu0 * (Sxf @ (dxf @ ezi))
            row.append({"omega": complex(omegan[j]), "data": _pack(g, ezi, hx, hy), "pol": TM})
        out.append(row)
    return out


# --- flux consumer: src/flux.jl:37-47 (TM, x̂ normal only) -----------------------
def flux_surface_integral_tm_x(g: Grid, data, center, width):
    """data is (Nx,Ny,3) [Ez,Hx,Hy].  x-index = the centre within dx/2 of center.x; when two
    centres tie (point exactly on a cell edge) the lower index is what the notebook got."""
    xs = xc(g)
    hits = np.nonzero(np.abs(xs - center[0]) <= dx(g) / 2 * (1 + 1e-9))[0]
    if len(hits) == 0:
        raise IndexError("no x-centre within dx/2")
    xi = int(hits[0])
    ys = yc(g)
    ysel = np.nonzero((ys >= center[1] - width) & (ys <= center[1] + width))[0]
    ez = (data[xi, ysel, 0] + data[xi + 1, ysel, 0]) / 2
    hy = data[xi, ysel, 2]
    return float(np.sum(-0.5 * np.real(ez * np.conj(hy))) * dy(g))


# --- closed-form coefficient view, used to check the CUDA stencil --
def stencil_coefficients(g: Grid, omega, ordering="fb"):
    """1-D off-diagonal coefficient arrays of the TM operator (without the ω²ε term):
        (A u)[ix,iy] = cxm[ix] u[ix-1] + cxp[ix] u[ix+1] + cym[iy] u[iy-1] + cyp[iy] u[iy+1]
                       - (cxm+cxp)[ix] u - (cym+cyp)[iy] u + ω²ε u
    'fb' = driven.jl:35 (Dxf·Dxb), 'bf' = modulation.jl:82 (Dxb·Dxf)."""
    _, mu0, _ = normalize_parameters(g)
    sxf, sxb, syf, syb = inv_sfactors(g, omega)
    ax, ay = 1 / dx(g), 1 / dy(g)

    def one(sf, sb, a):
        if ordering == "fb":
            cm = (sf * a) * (1 / mu0) * (sb * a)
            cp = (sf * a) * (1 / mu0) * (np.roll(sb, -1) * a)
        else:
            cp = (sb * a) * (1 / mu0) * (sf * a)
            cm = (sb * a) * (1 / mu0) * (np.roll(sf, 1) * a)
        return cm, cp

    cxm, cxp = one(sxf, sxb, ax)
    cym, cyp = one(syf, syb, ay)
    return cxm, cxp, cym, cyp
