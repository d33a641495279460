"""Synthetic permittivity maps for the benchmark configurations (BASELINE.json `configs`).
Deterministic (seed 0).  Host-side NumPy; only the resulting eps_r / src arrays go to the GPU."""
from __future__ import annotations

import math

import numpy as np

OMEGA_200THZ = 2 * math.pi * 200e12


def synthetic_tm_device(fdfd, Nx, Ny, dh=0.02, npml=15, seed=0, density=1.0 / 40.0, rows=None):
    """Vacuum + eps=12 straight waveguide (0.3 um wide, along x through the centre, runtests.jl:26) + a seeded
    set of eps in [2, 12.25] cylinders/boxes; x-normal line source at ix = npml + 10 (device.jl:103-104).
    dh = 0.02 um = lambda0/75 at 200 THz (notebook cell 14).
    rows=(y0, n): rasterise only global rows [y0, y0+n) (slab-sharded runs: every rank builds its own rows of the same
    map); returns (grid, omega, eps_rows, src_rows) with (Nx, n) arrays instead of a Device."""
    g = fdfd.Grid(dh, [npml, npml], [0.0, Nx * dh], [0.0, Ny * dh])
    assert g.N == (Nx, Ny)
    yall = fdfd.yc(g)
    if rows is not None:
        yall = yall[rows[0]:rows[0] + rows[1]]
    xall = fdfd.xc(g)
    xs, ys = xall[:, None], yall[None, :]
    eps = np.ones((Nx, len(yall)))
    rng = np.random.default_rng(seed)
    Lx, Ly = Nx * dh, Ny * dh
    nshape = max(4, int(Lx * Ly * density))
    for k in range(nshape):
        cx, cy = rng.uniform(0.1 * Lx, 0.9 * Lx), rng.uniform(0.1 * Ly, 0.9 * Ly)
        e = rng.uniform(2, 12.25)
        if k % 2 == 0:
            r = rng.uniform(0.3, 1.5)
            hx = hy = r
        else:
            wx, wy = rng.uniform(0.3, 3), rng.uniform(0.3, 3)
            hx, hy = wx / 2, wy / 2
        # only the bounding box of the shape is touched (the map has hundreds of shapes at 16384^2)
        i0, i1 = np.searchsorted(xall, cx - hx - dh), np.searchsorted(xall, cx + hx + dh)
        j0, j1 = np.searchsorted(yall, cy - hy - dh), np.searchsorted(yall, cy + hy + dh)
        if i0 >= i1 or j0 >= j1:
            continue
        sx, sy, sub = xs[i0:i1], ys[:, j0:j1], eps[i0:i1, j0:j1]
        if k % 2 == 0:
            sub[(sx - cx) ** 2 + (sy - cy) ** 2 <= r * r] = e
        else:
            sub[(np.abs(sx - cx) <= wx / 2) & (np.abs(sy - cy) <= wy / 2)] = e
    eps[:, np.abs(yall - Ly / 2) <= 0.15] = 12.0
    if rows is not None:
        src = np.zeros(eps.shape, dtype=np.complex128)
        src[npml + 10, :] = 1j
        return g, OMEGA_200THZ, eps.astype(np.complex128), src
    d = fdfd.Device(g, OMEGA_200THZ)
    d.eps_r = eps.astype(np.complex128)
    d.src[npml + 10, :] = 1j
    return d


def directional_coupler(fdfd, Nx=2000, Ny=1000, dh=0.0025, npml=15, gap=0.2, width=0.3, eps_core=12.0):
    """BASELINE config 2 (README figure, 4.5 x 2.5 um at 2000 x 1000): two parallel waveguides that approach
    over the central third; mode source on the upper guide.  No script exists in the reference (notebook cell 35)."""
    Lx, Ly = Nx * dh, Ny * dh
    g = fdfd.Grid(dh, [npml, npml], [0.0, Lx], [-Ly / 2, Ly / 2])
    d = fdfd.Device(g, OMEGA_200THZ)
    xs = fdfd.xc(g)[:, None]; ys = fdfd.yc(g)[None, :]
    far = 0.5
    t = np.clip((np.minimum(xs, Lx - xs) - 0.2 * Lx) / (0.15 * Lx), 0.0, 1.0)
    sep = far + (gap / 2 + width / 2 - far) * (3 * t ** 2 - 2 * t ** 3)  # smooth S-bend of the guide centres
    eps = np.ones((Nx, Ny))
    eps[np.abs(ys - sep) <= width / 2] = eps_core
    eps[np.abs(ys + sep) <= width / 2] = eps_core
    d.eps_r = eps.astype(np.complex128)
    fdfd.add_mode(d, fdfd.Mode(fdfd.TM, fdfd.XHAT, 3.0, fdfd.Point(0.35, far), 0.8))
    return d


def photonic_crystal_slab(fdfd, Nx=1024, Ny=1024, a=0.5, r_over_a=0.2, eps_rod=12.25, npml=15, nfreq=64):
    """BASELINE config 3: square lattice of rods, TE, nfreq frequencies uniformly in 2 pi [150, 250] THz."""
    dh = 0.02
    g = fdfd.Grid(dh, [npml, npml], [0.0, Nx * dh], [0.0, Ny * dh])
    ws = [2 * math.pi * f for f in np.linspace(150e12, 250e12, nfreq)]
    d = fdfd.Device(g, ws)
    xs = fdfd.xc(g)[:, None]; ys = fdfd.yc(g)[None, :]
    Lx, Ly = Nx * dh, Ny * dh
    inside = (xs > 0.2 * Lx) & (xs < 0.8 * Lx) & (ys > 0.1 * Ly) & (ys < 0.9 * Ly)
    fx = (xs / a) % 1.0 - 0.5; fy = (ys / a) % 1.0 - 0.5
    rods = (fx ** 2 + fy ** 2 <= r_over_a ** 2) & inside
    eps = np.ones((Nx, Ny)); eps[rods] = eps_rod
    d.eps_r = eps.astype(np.complex128)
    d.src[npml + 10, :] = 1j
    return d


def ring_resonator(fdfd, dh=0.01, R1=1.0, W=0.2, Wx=4.0, Wy=4.0, npml=15):
    """BASELINE config 4 / notebook cell 31: ring of eps 12.25 between radii R1-W and R1."""
    g = fdfd.Grid(dh, [npml, npml], [-Wx / 2, Wx / 2], [-Wy / 2, Wy / 2])
    d = fdfd.Device(g, OMEGA_200THZ)
    fdfd.setup_eps_r(d, [fdfd.Cylinder((0, 0), R1 - W, 1.0), fdfd.Cylinder((0, 0), R1, 12.25)])
    return d
