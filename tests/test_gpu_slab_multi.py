"""GPU tests of csrc/slab_multi.cu (slab-sharded modulated and eigenfrequency solves).

First run on hardware in round 2 (11/11 green).
All slabs live on ONE GPU (thread transport), like tests/test_gpu_slab.py.  Bars: true relative residual <= 1e-10, fields
within 1e-6 relative L2 of the single-GPU solve and of the oracle, eigenfrequencies within 1e-8 relative."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-6
EIG_TOL = 1e-8


def _mod_device(fdfd, L, dh, ns, sharedpml=True, Ny_pad=0.0):
    w, Om = 2 * math.pi * 1.939e14, 4.541e14
    a, q = 0.2202, 2.9263
    g = fdfd.Grid(dh, [15, 15], [0.0, L], [-1.28 - Ny_pad, 1.28 + Ny_pad])   # Ny = 128: divisible by 4 slabs x 2^levels
    d = fdfd.ModulatedDevice(g, w, Om, ns, sharedpml=sharedpml)
    fdfd.setup_eps_r(d, lambda x, y: -a / 2 <= y <= a / 2, 12.25)
    fdfd.setup_deps_r(d, lambda x, y: (1 <= x <= (L - 1)) and (-a / 2 <= y <= 0), lambda x, y: np.exp(1j * q * x))
    fdfd.add_mode(d, fdfd.Mode(fdfd.TM, fdfd.XHAT, 3.5, fdfd.Point(0.2, 0), 4 * a))
    return d


@pytest.mark.parametrize("nslabs", [1, 2, 4])
@pytest.mark.parametrize("sharedpml", [True, False])
def test_modulated_slabs_vs_single_gpu(fdfd, nslabs, sharedpml):
    from fdfd_jl_b200 import slab
    d = _mod_device(fdfd, 5.0, 0.02, 1, sharedpml)
    ref = fdfd.solve(d)[0]                      # also launches the mode source into d.src
    fs, infos = slab.solve_modulated_slabs_threads(d, nslabs)
    assert len(fs) == 3
    for i in infos:
        assert i["flag"] == 0 and i["relres"] <= 1e-10 and i["iters"] == infos[0]["iters"]
    scale = max(np.linalg.norm(f.data) for f in ref)
    for j in range(3):
        assert abs(fs[j].omega - ref[j].omega) <= 1e-12 * abs(ref[j].omega)
        assert np.linalg.norm(fs[j].data - ref[j].data) / scale <= FIELD_TOL


def test_modulated_slabs_vs_oracle(fdfd):
    from fdfd_jl_b200 import slab
    from oracle import fdfd_oracle as O
    d = _mod_device(fdfd, 3.0, 0.02, 1)
    fdfd.solve(d)                               # mode source
    fs, _ = slab.solve_modulated_slabs_threads(d, 2)
    g = d.grid
    go = O.Grid2D(0.02, list(g.Npml), [g.bounds[0][0], g.bounds[1][0]], [g.bounds[0][1], g.bounds[1][1]])
    do = O.ModulatedDevice(go, [d.omega[0]], Omega=d.Omega, nsidebands=1, sharedpml=True)
    do.eps_r[:] = d.eps_r; do.deps_r[:] = d.deps_r; do.src[:] = d.src
    fo = O.solve_modulated(do)[0]
    scale = max(np.linalg.norm(f["data"]) for f in fo)
    for j in range(3):
        assert np.linalg.norm(fs[j].data - fo[j]["data"]) / scale <= FIELD_TOL


def test_modulated_slab_no_sidebands_is_bf_driven(fdfd):
    from fdfd_jl_b200 import slab
    d = _mod_device(fdfd, 3.0, 0.02, 0)
    ref = fdfd.solve(d)[0]
    fs, infos = slab.solve_modulated_slabs_threads(d, 2)
    assert len(fs) == 1 and infos[0]["flag"] == 0
    assert np.linalg.norm(fs[0].data - ref[0].data) / np.linalg.norm(ref[0].data) <= FIELD_TOL


def _match(got, ref):
    ref = list(ref)
    worst = 0.0
    for z in got:
        k = int(np.argmin([abs(z - r) for r in ref]))
        worst = max(worst, abs(z - ref[k]) / abs(ref[k]))
        ref.pop(k)
    return worst


@pytest.mark.parametrize("nslabs", [1, 2, 4])
def test_eigenfrequency_slabs_ring(fdfd, nslabs):
    """notebook cell 31 ring at 256^2 (power-of-two rows so that 4 slabs keep a 3-level hierarchy)"""
    from fdfd_jl_b200 import slab
    g = fdfd.Grid(4.0 / 256, [15, 15], [-2.0, 2.0], [-2.0, 2.0])
    d = fdfd.Device(g, 2 * math.pi * 200e12)
    fdfd.setup_eps_r(d, [fdfd.Cylinder((0, 0), 0.8, 1.0), fdfd.Cylinder((0, 0), 1.0, 12.25)])
    nev = 4
    om_ref, f_ref = fdfd.eigenfrequency(d, fdfd.TM, nev + 2, which="LM")
    om, fields, infos = slab.eigenfrequency_slabs_threads(d, nev, nslabs, which="LM")
    assert _match(om, om_ref) <= EIG_TOL
    # eigen-pair residual with the assembled reference operator (eigen.jl:84): A = Teps^-1 (Dxf Dxb + Dyf Dyb)
    from oracle import fdfd_oracle as O
    go = O.Grid2D(4.0 / 256, [15, 15], [-2.0, 2.0], [-2.0, 2.0])
    do = O.Device(go, [d.omega[0]]); do.eps_r[:] = d.eps_r
    A, sigma, aux = O.eigen_matrix(do, O.TM)
    eps0, mu0, _ = O.normalize_parameters(go)
    for i in range(nev):
        x = fields[i].data[:, :, 0].ravel(order="F")
        lam = -(om[i] ** 2) * mu0 * eps0
        assert np.linalg.norm(A @ x - lam * x) / np.linalg.norm(lam * x) < 1e-6
        assert abs(np.linalg.norm(x) - 1) < 1e-8
