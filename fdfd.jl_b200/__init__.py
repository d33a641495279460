"""fdfd.jl_b200 -- host-side mirror of the FDFD.jl API for the assembly + linear-solve hot path.

Same names, argument meaning and error behaviour as the reference's Julia functions (file:line cited per
function), calling the sm_100a CUDA library through its C ABI (include/fdfd_b200.h).  The Julia `ccall`
wrapper a maintainer would drop into FDFD.jl is in `julia/FDFDB200.jl` + INTEGRATION.md; Julia is not
installed in the build image, so this Python mirror is what the tests drive.

The directory name contains a dot, so import it through the repo-root shim:  `import fdfd_jl_b200 as fdfd`.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field as _f

import numpy as np

from . import _lib
from ._lib import (TM, TE, Context, FdfdError, default_context, default_opts, as_c128, ptr, check, lib,
                   GridT, Info, SolveOpts)

# src/types.jl:6-9
EPS0 = 8.85418782e-12
MU0 = 1.25663706e-6
C0 = math.sqrt(1 / EPS0 / MU0)
ETA0 = math.sqrt(MU0 / EPS0)
DEFAULT_L0 = 1e-6
XHAT, YHAT = 1, 2  # Direction x̂, ŷ (src/types.jl:18)


@dataclass(frozen=True)
class Point:  # src/types.jl:23-26
    x: float
    y: float


def _pt(p):
    return p if isinstance(p, Point) else Point(float(p[0]), float(p[1]))


class Grid:
    """Grid{2} (src/grid.jl:7-33).  Grid(dh, Npml, xrange, yrange; L0): N = round.(L/dh) (ties to even)."""

    def __init__(self, dh, Npml, xrange, yrange, L0=DEFAULT_L0):
        xr, yr = np.ravel(xrange).astype(float), np.ravel(yrange).astype(float)
        np_ = np.ravel(Npml).astype(int)
        self.L = (float(xr[1] - xr[0]), float(yr[1] - yr[0]))
        self.L0 = float(L0)
        self.N = (int(round(self.L[0] / dh)), int(round(self.L[1] / dh)))
        self.Npml = (int(np_[0]), int(np_[1]))
        self.bounds = ((float(xr[0]), float(yr[0])), (float(xr[1]), float(yr[1])))

    def size(self):
        return self.N

    def __len__(self):
        return self.N[0] * self.N[1]

    def as_c(self) -> GridT:
        return GridT(self.N[0], self.N[1], self.Npml[0], self.Npml[1], self.bounds[0][0], self.bounds[1][0],
                     self.bounds[0][1], self.bounds[1][1], self.L0)


def dx(g: Grid):  # src/grid.jl:68-70
    return (g.bounds[1][0] - g.bounds[0][0]) / g.N[0]


def dy(g: Grid):  # src/grid.jl:72-74
    return (g.bounds[1][1] - g.bounds[0][1]) / g.N[1]


def xc(g: Grid):  # src/grid.jl:76-78
    return g.bounds[0][0] + dx(g) * (0.5 + np.arange(g.N[0]))


def yc(g: Grid):  # src/grid.jl:80-82
    return g.bounds[0][1] + dy(g) * (0.5 + np.arange(g.N[1]))


def xe(g: Grid):  # src/grid.jl:89-91
    return g.bounds[0][0] + dx(g) * np.arange(g.N[0] + 1)


def ye(g: Grid):  # src/grid.jl:93-95
    return g.bounds[0][1] + dy(g) * np.arange(g.N[1] + 1)


def x2ind(g: Grid, x):  # src/grid.jl:112-117, 1-based like Julia
    ind = int(round((x - g.bounds[0][0]) / g.L[0] * g.N[0]) + 1)
    return min(max(ind, 1), g.N[0])


def y2ind(g: Grid, y):  # src/grid.jl:120-125
    ind = int(round((y - g.bounds[0][1]) / g.L[1] * g.N[1]) + 1)
    return min(max(ind, 1), g.N[1])


def coord2ind(g: Grid, pt):  # src/grid.jl:103-109
    pt = _pt(pt)
    return x2ind(g, pt.x), y2ind(g, pt.y)


def normalize_parameters(g: Grid):  # src/device.jl:40
    return EPS0 * g.L0, MU0 * g.L0, C0 / g.L0


@dataclass
class Mode:  # src/device.jl:5-11
    pol: int
    dir: int
    neff: float
    pt: object
    width: float


# GeometryPrimitives stand-ins used by the reference's examples (Box / Cylinder in the xy-plane)
@dataclass
class Box:
    center: tuple
    size: tuple
    data: object

    def contains(self, x, y):
        return (np.abs(x - self.center[0]) <= self.size[0] / 2) & (np.abs(y - self.center[1]) <= self.size[1] / 2)


@dataclass
class Cylinder:
    center: tuple
    radius: float
    data: object

    def contains(self, x, y):
        return (x - self.center[0]) ** 2 + (y - self.center[1]) ** 2 <= self.radius ** 2


class Device:
    """Device{2} (src/device.jl:19-35): eps_r ones, src zeros, omega vector, modes."""

    def __init__(self, grid: Grid, omega):
        self.grid = grid
        self.omega = [float(omega)] if np.isscalar(omega) else [float(w) for w in omega]
        self.eps_r = np.ones(grid.N, dtype=np.complex128)
        self.src = np.zeros(grid.N, dtype=np.complex128)
        self.modes = []


class ModulatedDevice(Device):
    """ModulatedDevice (src/solver/modulation.jl:4-26)."""

    def __init__(self, grid: Grid, omega, Omega, nsidebands, sharedpml=True):
        super().__init__(grid, omega)
        self.Omega = float(Omega)
        self.nsidebands = int(nsidebands)
        self.sharedpml = bool(sharedpml)
        self.deps_r = np.zeros(grid.N, dtype=np.complex128)


def _mask_values(pixels, g: Grid, region, value):  # src/device.jl:63-83
    X, Y = np.meshgrid(xc(g), yc(g), indexing="ij")
    mask = np.asarray(np.vectorize(region)(X, Y), dtype=bool)
    if callable(value):
        pixels[mask] = np.vectorize(value)(X, Y)[mask]
    else:
        pixels[mask] = value


def _compose_shapes(pixels, g: Grid, shapes):  # src/device.jl:47-61 (first shape containing the pixel wins)
    X, Y = np.meshgrid(xc(g), yc(g), indexing="ij")
    done = np.zeros(pixels.shape, dtype=bool)
    for sh in shapes:
        mask = sh.contains(X, Y) & ~done
        pixels[mask] = np.vectorize(sh.data)(X, Y)[mask] if callable(sh.data) else sh.data
        done |= mask


def setup_eps_r(d: Device, *args):
    """setup_ϵᵣ!(d, shapes) / setup_ϵᵣ!(d, region, value)  (src/device.jl:86-89)."""
    if len(args) == 1:
        _compose_shapes(d.eps_r, d.grid, args[0])
    else:
        _mask_values(d.eps_r, d.grid, args[0], args[1])


def setup_deps_r(d: ModulatedDevice, *args):
    """setup_Δϵᵣ!  (src/solver/modulation.jl:29-32)."""
    if len(args) == 1:
        _compose_shapes(d.deps_r, d.grid, args[0])
    else:
        _mask_values(d.deps_r, d.grid, args[0], args[1])


def setup_src(d: Device, *args):
    """setup_src!(d, region, value) | setup_src!(d, pt) | setup_src!(d, pt, srcnormal)  (src/device.jl:92-108)."""
    if len(args) == 2 and callable(args[0]):
        return _mask_values(d.src, d.grid, args[0], args[1])
    ix, iy = coord2ind(d.grid, args[0])
    if len(args) == 1:
        d.src[ix - 1, iy - 1] = 1j
    elif args[1] == XHAT:
        d.src[ix - 1, :] = 1j
    elif args[1] == YHAT:
        d.src[:, iy - 1] = 1j


def add_mode(d: Device, mode: Mode):  # src/device.jl:113-115
    d.modes.append(mode)


def _grid_average(a, axis):  # src/grid.jl:157-162
    return (a + np.roll(a, 1, axis=axis)) / 2


def eigenmode_1d(eps_slice, h, L0, omega, pol, neff, nev=1):
    """1-D slice mode (src/solver/eigen.jl:6-29 on the Npml=0 grid of src/device.jl:144).  <=~100 unknowns:
    stays on the host, dense eigendecomposition, eigenvalues nearest sigma = (ω/c₀·neff)²."""
    n = len(eps_slice)
    eps0, mu0, c0 = EPS0 * L0, MU0 * L0, C0 / L0
    c = 1 / h
    i = np.arange(n)
    dxf = np.zeros((n, n)); dxf[i, i] = -c; dxf[i, (i + 1) % n] = c
    dxb = np.zeros((n, n)); dxb[i, i] = c; dxb[i, (i - 1) % n] = -c
    if pol == TM:
        A = omega ** 2 * mu0 * np.diag(eps0 * eps_slice) + dxf @ dxb
    else:
        A = omega ** 2 * mu0 * np.diag(eps0 * eps_slice) + np.diag(eps0 * eps_slice) @ dxf @ np.diag(1 / (eps0 * _grid_average(eps_slice, 0))) @ dxb
    sigma = (omega / c0 * neff) ** 2
    w, v = np.linalg.eig(A)
    order = np.argsort(np.abs(w - sigma))[:nev]
    return np.sqrt(w[order].astype(np.complex128)), v[:, order]


def get_modes(d: Device, pol, omega, neff, nmodes, pt, slicenormal, slicewidth):
    """src/device.jl:124-149.  Returns (beta, vectors, ix0, iy0) with 0-based index (arrays)."""
    g = d.grid
    ix, iy = coord2ind(g, pt)
    srcpoints = int(round(slicewidth / (dy(g) if slicenormal == XHAT else dx(g))))
    if srcpoints % 2 == 0:
        srcpoints += 1
    M = (srcpoints - 1) // 2
    if slicenormal == XHAT:
        iy = iy + np.arange(-M, M + 1); h = dy(g)
    else:
        ix = ix + np.arange(-M, M + 1); h = dx(g)
    # Julia throws a BoundsError when the slice leaves the grid (device.jl:133-146); NumPy would wrap negative indices silently
    Nx, Ny = g.N
    if np.min(ix) < 1 or np.max(ix) > Nx or np.min(iy) < 1 or np.max(iy) > Ny:
        raise IndexError(f"mode slice x in [{np.min(ix)}, {np.max(ix)}], y in [{np.min(iy)}, {np.max(iy)}] leaves the {Nx} x {Ny} grid")
    eps_slice = d.eps_r[ix - 1, iy - 1]
    beta, vec = eigenmode_1d(eps_slice, h, g.L0, omega, pol, neff, nmodes)
    return beta, vec, ix - 1, iy - 1


def setup_mode(d: Device, pol, omega, neff, pt, srcnormal, srcwidth):  # src/device.jl:118-121
    _, vec, ix, iy = get_modes(d, pol, omega, neff, 1, pt, srcnormal, srcwidth)
    v = np.abs(vec[:, 0])
    d.src[ix, iy] += v / np.linalg.norm(v)


def _apply_modes(d: Device, omega):  # src/solver/driven.jl:15-19 (modes always launched as TM)
    if d.modes:
        d.src = np.zeros(d.grid.N, dtype=np.complex128)
    for m in d.modes:
        setup_mode(d, TM, omega, m.neff, m.pt, m.dir, m.width)


class Field:
    """FieldTM / FieldTE (src/data.jl:50-78): data is (Nx,Ny,3), components by name."""
    components = ()

    def __init__(self, grid, omega, data, info=None):
        self.grid, self.omega, self.data, self.info = grid, complex(omega), data, info

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.data[:, :, self.components.index(key)]
        return self.data[key]

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)


class FieldTM(Field):
    components = ("Ez", "Hx", "Hy")


class FieldTE(Field):
    components = ("Hz", "Ex", "Ey")


def _opts(kw) -> SolveOpts:
    return kw.pop("opts", None) or default_opts(**kw)


def solve(d, pol=TM, ctx: Context = None, **kw):
    """solve(d::Device, pol=TM) (src/solver/driven.jl:4-59) and solve(d::ModulatedDevice)
    (src/solver/modulation.jl:35-119).  Returns a Field, or a list over ω (driven.jl:57-58); for a
    ModulatedDevice a list [iω][sideband]."""
    if isinstance(d, ModulatedDevice):
        return _solve_modulated(d, ctx, **kw)
    ctx = ctx or default_context()
    o = _opts(kw)
    g = d.grid
    Nx, Ny = g.N
    gc = g.as_c()
    nw = len(d.omega)
    eps = as_c128(d.eps_r, (Nx, Ny))
    srcs = np.empty((Nx, Ny, nw), dtype=np.complex128, order="F")
    for i, omega in enumerate(d.omega):
        _apply_modes(d, omega)  # mode source depends on ω; host-side, tiny (device.jl:118-149)
        srcs[:, :, i] = d.src
    fields = np.empty((Nx, Ny, 3, nw), dtype=np.complex128, order="F")
    infos = (Info * nw)()
    w = (C.c_double * nw)(*d.omega)
    # one call for the whole sweep: the library solves up to `concurrency` frequencies at the same time
    code = lib().fdfd_solve_driven(ctx.handle, C.byref(gc), pol, nw, w, ptr(eps), ptr(srcs), 1, C.byref(o), ptr(fields), infos)
    check(code, ctx.handle)
    cls = FieldTM if pol == TM else FieldTE
    out = [cls(g, d.omega[i], fields[:, :, :, i], infos[i].asdict()) for i in range(nw)]
    return out[0] if len(out) == 1 else out


def _solve_modulated(d: ModulatedDevice, ctx=None, **kw):
    ctx = ctx or default_context()
    o = _opts(kw)
    g = d.grid
    Nx, Ny = g.N
    gc = g.as_c()
    nf = 2 * d.nsidebands + 1
    out = []
    for omega in d.omega:
        _apply_modes(d, omega)
        eps = as_c128(d.eps_r, (Nx, Ny)); deps = as_c128(d.deps_r, (Nx, Ny)); src = as_c128(d.src, (Nx, Ny))
        fields = np.empty((Nx, Ny, 3, nf), dtype=np.complex128, order="F")
        info = Info()
        code = lib().fdfd_solve_modulated(ctx.handle, C.byref(gc), omega, d.Omega, d.nsidebands, int(d.sharedpml),
                                          ptr(eps), ptr(deps), ptr(src), C.byref(o), ptr(fields), C.byref(info))
        check(code, ctx.handle)
        omegan = omega + d.Omega * np.arange(-d.nsidebands, d.nsidebands + 1)
        out.append([FieldTM(g, omegan[j], fields[:, :, :, j], info.asdict()) for j in range(nf)])
    return out


def _csc_arrays(A):
    """(n, colptr, rowval, nzval) of a square SciPy sparse matrix / (colptr, rowval, nzval) triple, Int64 / ComplexF64, 0-based."""
    if isinstance(A, tuple):
        colptr, rowval, nzval = A
        n = len(colptr) - 1
    else:
        A = A.tocsc()
        if A.shape[0] != A.shape[1]:
            raise ValueError("dolinearsolve needs a square matrix")
        n, colptr, rowval, nzval = A.shape[0], A.indptr, A.indices, A.data
    return (n, np.ascontiguousarray(colptr, dtype=np.int64), np.ascontiguousarray(rowval, dtype=np.int64),
            np.ascontiguousarray(nzval, dtype=np.complex128))


def dolinearsolve(A, b, matrixsym=None, index_base=0, grid: Grid = None, omega=None, ctx: Context = None, return_info=False, **kw):
    """dolinearsolve(A::SparseMatrixCSC, b, matrixsym) -> x (src/solver/solver.jl:4-41) for callers that assemble their own matrix
    (nonlinear.jl:69,97,120; eigen.jl:32-66).  A: SciPy sparse matrix or a (colptr, rowval, nzval) CSC triple with the given
    index_base; `matrixsym` is accepted and ignored, as in the reference (solver.jl:29).  BiCGSTAB + Jacobi on the GPU
    (fdfd_dolinearsolve_csc); options: tol, maxit, check_every, use_graph, verbose.  With `grid` and `omega` (the grid A was
    assembled on) the library checks whether A is the TM operator of that grid for some permittivity -- the first solve and the
    Born steps of nonlinear.jl are -- and then runs the multigrid-preconditioned solver (fdfd_dolinearsolve_csc_grid;
    info["mg_levels"] > 0), else the generic path."""
    n, colptr, rowval, nzval = _csc_arrays(A)
    bb = np.ascontiguousarray(np.asarray(b, dtype=np.complex128).ravel(order="F"))
    if bb.size != n:
        raise ValueError(f"b has {bb.size} entries, A is {n} x {n}")
    ctx = ctx or default_context()
    o = _opts(kw)
    x = np.empty(n, dtype=np.complex128)
    info = Info()
    if grid is not None:
        if omega is None:
            raise ValueError("dolinearsolve: grid given without omega")
        gc = grid.as_c()
        code = lib().fdfd_dolinearsolve_csc_grid(ctx.handle, C.byref(gc), float(omega), n, ptr(colptr), ptr(rowval), ptr(nzval),
                                                 int(index_base), ptr(bb), C.byref(o), ptr(x), C.byref(info))
    else:
        code = lib().fdfd_dolinearsolve_csc(ctx.handle, n, ptr(colptr), ptr(rowval), ptr(nzval), int(index_base), ptr(bb), C.byref(o),
                                            ptr(x), C.byref(info))
    check(code, ctx.handle)
    return (x, info.asdict()) if return_info else x


def _sell_spmv_host(A, x, index_base=0, rowsum=None):
    """host-only test hook: y = A x through the library's CSC -> SELL-32 transposition (no GPU) -> (y, dinv, padded_entries);
    rowsum: optional complex128 (n) output array receiving A 1"""
    n, colptr, rowval, nzval = _csc_arrays(A)
    xx = np.ascontiguousarray(x, dtype=np.complex128)
    y = np.empty(n, dtype=np.complex128); dinv = np.empty(n, dtype=np.complex128)
    pad = C.c_int64(0)
    rs = None if rowsum is None else rowsum
    code = lib().fdfd_debug_sell_spmv(n, ptr(colptr), ptr(rowval), ptr(nzval), int(index_base), ptr(xx), ptr(y), ptr(dinv), ptr(rs), C.byref(pad))
    if code != 0:
        raise FdfdError(code, "fdfd_debug_sell_spmv: bad CSC arrays")
    return y, dinv, pad.value


def eigenfrequency(d: Device, pol, nev, which="LM", ncv=0, ctx: Context = None, **kw):
    """eigenfrequency(d, pol, neigenvalues; which=:LM) (src/solver/eigen.jl:69-115) -> (ω, fields)."""
    ctx = ctx or default_context()
    o = _opts(kw)
    g = d.grid
    Nx, Ny = g.N
    gc = g.as_c()
    eps = as_c128(d.eps_r, (Nx, Ny))
    om = np.empty(nev, dtype=np.complex128)
    fields = np.empty((Nx, Ny, 3, nev), dtype=np.complex128, order="F")
    info = Info()
    code = lib().fdfd_eigenfrequency(ctx.handle, C.byref(gc), pol, d.omega[0], nev, _lib.WHICH[which], ncv, ptr(eps),
                                     C.byref(o), ptr(om), ptr(fields), C.byref(info))
    check(code, ctx.handle)
    cls = FieldTM if pol == TM else FieldTE
    return om, [cls(g, om[i], fields[:, :, :, i], info.asdict()) for i in range(nev)]


# ---- consumers of the returned fields (src/flux.jl), host-side: they read O(Ny) values ------------------
def probe_field(field: Field, component, pt):  # src/flux.jl:6-12
    pt = _pt(pt)
    g = field.grid
    xi = np.nonzero(np.abs(xc(g) - pt.x) <= dx(g) / 2)[0][0]
    yi = np.nonzero(np.abs(yc(g) - pt.y) <= dy(g) / 2)[0][0]
    if component not in field.components:
        raise ValueError(f"{component} is invalid for this polarization. Valid options are {field.components}")
    return field[component][xi, yi]


def flux_surface_integral(field: Field, center, width, normal):
    """src/flux.jl:37-47 (TM, x̂ normal; the only working branch of the reference)."""
    center = _pt(center)
    if normal != XHAT:
        raise NotImplementedError("ŷ normal calculation not yet implemented")  # flux.jl:58-65
    if not isinstance(field, FieldTM):
        raise NotImplementedError("TE flux is broken in the reference (flux.jl:48-55)")
    g = field.grid
    hits = np.nonzero(np.abs(xc(g) - center.x) <= dx(g) / 2 * (1 + 1e-9))[0]
    if len(hits) == 0:
        raise IndexError("no x-centre within dx/2 of center.x")
    xi = int(hits[0])
    ys = yc(g)
    sel = np.nonzero((ys >= center.y - width) & (ys <= center.y + width))[0]
    ez = (field.data[xi, sel, 0] + field.data[xi + 1, sel, 0]) / 2
    hy = field.data[xi, sel, 2]
    return float(np.sum(-0.5 * np.real(ez * np.conj(hy))) * dy(g))


# ---- kernel-level hooks (parity tests / benchmarks) --------------------------------------------------------
def sfactors(g: Grid, omega, ctx=None):
    """create_sfactor for (x̂,F), (x̂,B), (ŷ,F), (ŷ,B)  (src/pml.jl:1-31)."""
    ctx = ctx or default_context()
    Nx, Ny = g.N
    outs = [np.empty(n, dtype=np.complex128) for n in (Nx, Nx, Ny, Ny)]
    gc = g.as_c()
    check(lib().fdfd_sfactors(ctx.handle, C.byref(gc), omega, *[ptr(a) for a in outs]), ctx.handle)
    return outs


def assemble_derivative(g: Grid, omega, which, stretched=True, fmt=_lib.CSR, index_base=0, ctx=None):
    ctx = ctx or default_context()
    N = len(g)
    p = np.empty(N + 1, dtype=np.int64); ind = np.empty(2 * N, dtype=np.int64); val = np.empty(2 * N, dtype=np.complex128)
    gc = g.as_c()
    check(lib().fdfd_assemble_derivative(ctx.handle, C.byref(gc), omega, which, int(stretched), fmt, index_base,
                                         ptr(p), ptr(ind), ptr(val)), ctx.handle)
    return p, ind, val


def assemble_system(g: Grid, pol, omega, eps_r, ordering=_lib.ORDER_FB, fmt=_lib.CSR, index_base=0, ctx=None):
    ctx = ctx or default_context()
    N = len(g)
    p = np.empty(N + 1, dtype=np.int64); ind = np.empty(5 * N, dtype=np.int64); val = np.empty(5 * N, dtype=np.complex128)
    eps = as_c128(eps_r, g.N)
    gc = g.as_c()
    check(lib().fdfd_assemble_system(ctx.handle, C.byref(gc), pol, ordering, omega, ptr(eps), fmt, index_base,
                                     ptr(p), ptr(ind), ptr(val)), ctx.handle)
    return p, ind, val


def apply_operator(g: Grid, pol, omega, eps_r, x, ordering=_lib.ORDER_FB, ctx=None):
    ctx = ctx or default_context()
    eps = as_c128(eps_r, g.N); xx = as_c128(x, g.N)
    y = np.empty(g.N, dtype=np.complex128, order="F")
    gc = g.as_c()
    check(lib().fdfd_apply_operator(ctx.handle, C.byref(gc), pol, ordering, omega, ptr(eps), ptr(xx), ptr(y)), ctx.handle)
    return y


def apply_operator_batched(g: Grid, pol, omega, eps_r, X, ordering=_lib.ORDER_FB, ctx=None):
    """Y[:, :, b] = A X[:, :, b] for right-hand sides sharing one operator: the stencil reads the coefficients once per point for
    all of them ((32 B + 16) / B bytes per point and right-hand side instead of 48).  X: (Nx, Ny, B)."""
    X = np.asarray(X)
    if X.ndim != 3 or X.shape[:2] != tuple(g.N):
        raise ValueError("X must have shape (Nx, Ny, B)")
    B = X.shape[2]
    ctx = ctx or default_context()
    eps = as_c128(eps_r, g.N)
    xx = np.asfortranarray(X, dtype=np.complex128)
    Y = np.empty(tuple(g.N) + (B,), dtype=np.complex128, order="F")
    gc = g.as_c()
    check(lib().fdfd_apply_operator_batched(ctx.handle, C.byref(gc), pol, ordering, omega, ptr(eps), B, ptr(xx), ptr(Y)), ctx.handle)
    return Y


def rasterize(g: Grid, shapes, eps_r=None, ctx=None):
    """setup_ϵᵣ!(d, shapes) on the GPU for Box / Cylinder shapes with constant data (src/device.jl:47-61)."""
    ctx = ctx or default_context()
    eps = np.asfortranarray(np.ones(g.N, dtype=np.complex128) if eps_r is None else np.array(eps_r, dtype=np.complex128))
    rows = []
    for sh in shapes:
        v = complex(sh.data)
        if isinstance(sh, Box):
            rows.append([0.0, sh.center[0], sh.center[1], min(sh.size[0], 1e300), min(sh.size[1], 1e300), v.real, v.imag])
        elif isinstance(sh, Cylinder):
            rows.append([1.0, sh.center[0], sh.center[1], sh.radius, 0.0, v.real, v.imag])
        else:
            raise TypeError("rasterize handles Box and Cylinder")
    arr = np.ascontiguousarray(np.array(rows, dtype=np.float64).reshape(-1, 7))
    gc = g.as_c()
    check(lib().fdfd_rasterize(ctx.handle, C.byref(gc), len(rows), ptr(arr) if len(rows) else None, ptr(eps)), ctx.handle)
    return eps


class Problem:
    """Resident problem: eps_r, coefficients and the multigrid hierarchy stay in HBM between solves."""

    def __init__(self, g: Grid, pol, omega, eps_r, ordering=_lib.ORDER_FB, ctx=None, **kw):
        self.ctx = ctx or default_context()
        self.grid, self.pol, self.omega = g, pol, omega
        o = _opts(kw)
        self._h = C.c_void_p()
        gc = g.as_c()
        eps = ptr(eps_r) if isinstance(eps_r, (int, np.integer)) else ptr(as_c128(eps_r, g.N))
        check(lib().fdfd_problem_create(self.ctx.handle, C.byref(gc), pol, ordering, omega, eps, C.byref(o), C.byref(self._h)), self.ctx.handle)

    def set_source(self, src):
        s = src if isinstance(src, (int, np.integer)) else as_c128(src, self.grid.N)
        check(lib().fdfd_problem_set_source(self._h, ptr(s)), self.ctx.handle)

    def set_rhs(self, b):
        s = b if isinstance(b, (int, np.integer)) else as_c128(b, self.grid.N)
        check(lib().fdfd_problem_set_rhs(self._h, ptr(s)), self.ctx.handle)

    def solve(self):
        info = Info()
        check(lib().fdfd_problem_solve(self._h, C.byref(info)), self.ctx.handle)
        return info.asdict()

    def solution(self, out=None):
        x = np.empty(self.grid.N, dtype=np.complex128, order="F") if out is None else out
        check(lib().fdfd_problem_get_solution(self._h, ptr(x)), self.ctx.handle)
        return x

    def fields(self, forward_h=False, out=None):
        f = np.empty(self.grid.N + (3,), dtype=np.complex128, order="F") if out is None else out
        check(lib().fdfd_problem_get_fields(self._h, int(forward_h), ptr(f)), self.ctx.handle)
        return f

    def bench_apply(self, nrep=50):
        ms = C.c_double()
        check(lib().fdfd_problem_bench_apply(self._h, nrep, C.byref(ms)), self.ctx.handle)
        return ms.value

    def bench_apply_batched(self, nrhs, nrep=50):
        """ms per launch of the batched stencil: `nrhs` (1, 2, 4 or 8) right-hand sides sharing this problem's operator"""
        ms = C.c_double()
        check(lib().fdfd_problem_bench_apply_batched(self._h, int(nrhs), nrep, C.byref(ms)), self.ctx.handle)
        return ms.value

    def bench_mg(self, kind, nrep=50):
        """ms per launch of one level-0 multigrid kernel (0 smoother+correction, 1 restriction, 2 zero-guess sweep, 4 cycle)"""
        ms = C.c_double()
        check(lib().fdfd_problem_bench_mg(self._h, int(kind), nrep, C.byref(ms)), self.ctx.handle)
        return ms.value

    def flux_x(self, center, width, forward_h=False):
        """flux_surface_integral(field, center, width, x̂) (src/flux.jl:37-47) evaluated on the device from the resident Ez"""
        center = _pt(center)
        out = C.c_double()
        w = 1e300 if np.isinf(width) else float(width)
        check(lib().fdfd_problem_flux_x(self._h, center.x, center.y, w, int(forward_h), C.byref(out)), self.ctx.handle)
        return out.value

    def history(self, nmax=100000):
        """relative (recurrence) residual per iteration of the last solve"""
        out = np.empty(nmax, dtype=np.float64)
        n = C.c_int32()
        check(lib().fdfd_problem_get_history(self._h, ptr(out), nmax, C.byref(n)), self.ctx.handle)
        return out[:n.value].copy()

    def ml_cycles(self):
        """multigrid cycles started on levels 0..3 by the last FDFD_SOLVER_MLKRYLOV solve"""
        out = (C.c_int64 * 4)()
        check(lib().fdfd_problem_ml_cycles(self._h, out), self.ctx.handle)
        return list(out)

    def precond(self, v):
        vin = as_c128(v, self.grid.N)
        out = np.empty(self.grid.N, dtype=np.complex128, order="F")
        check(lib().fdfd_problem_precond(self._h, ptr(vin), ptr(out)), self.ctx.handle)
        return out

    def close(self):
        if self._h:
            lib().fdfd_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
