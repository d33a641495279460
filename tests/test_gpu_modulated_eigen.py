"""GPU parity for the modulated MF-FDFD solve (modulation.jl:35-119) and eigenfrequency (eigen.jl:69-115)."""
import math

import numpy as np
import pytest

from oracle import fdfd_oracle as O

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-6
EIG_TOL = 1e-8


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def _mod_device(fdfd, L, dh, ns, sharedpml=True):
    w, Om = 2 * math.pi * 1.939e14, 4.541e14
    a, q = 0.2202, 2.9263
    gargs = (dh, [15, 15], [0.0, L], [-1.0, 1.0])
    g, go = fdfd.Grid(*gargs), O.Grid2D(*gargs)
    d = fdfd.ModulatedDevice(g, w, Om, ns, sharedpml=sharedpml)
    fdfd.setup_eps_r(d, lambda x, y: -a / 2 <= y <= a / 2, 12.25)
    fdfd.setup_deps_r(d, lambda x, y: (1 <= x <= (L - 1)) and (-a / 2 <= y <= 0), lambda x, y: np.exp(1j * q * x))
    fdfd.add_mode(d, fdfd.Mode(fdfd.TM, fdfd.XHAT, 3.5, fdfd.Point(0.2, 0), 4 * a))
    do = O.ModulatedDevice(go, [w], Omega=Om, nsidebands=ns, sharedpml=sharedpml)
    do.eps_r[:] = d.eps_r
    do.deps_r[:] = d.deps_r
    do.modes.append(O.Mode(O.TM, O.X, 3.5, (0.2, 0), 4 * a))
    return d, do


@pytest.mark.parametrize("sharedpml", [True, False])
def test_modulated_waveguide_vs_oracle(fdfd, sharedpml):
    """test/runtests.jl:40-63 geometry (L=5) at dh=0.02 -> 250x100x3 unknowns."""
    d, do = _mod_device(fdfd, 5.0, 0.02, 1, sharedpml)
    fs = fdfd.solve(d)[0]
    fo = O.solve_modulated(do)[0]
    assert len(fs) == 3
    ref_norm = max(np.linalg.norm(f["data"]) for f in fo)
    for j in range(3):
        assert fs[j].info["flag"] == 0 and fs[j].info["relres"] <= 1e-10
        assert abs(fs[j].omega - fo[j]["omega"]) <= 1e-12 * abs(fo[j]["omega"])
        # sidebands are compared on the scale of the strongest one (the weak ones are pure coupling products)
        assert np.linalg.norm(fs[j].data - fo[j]["data"]) / ref_norm <= FIELD_TOL


def test_modulated_no_sidebands_is_bf_driven(fdfd):
    d, do = _mod_device(fdfd, 3.0, 0.02, 0)
    fs = fdfd.solve(d)[0]
    fo = O.solve_modulated(do)[0]
    assert len(fs) == 1 and rel(fs[0].data, fo[0]["data"]) <= FIELD_TOL


def test_notebook_photon_numbers_on_gpu(fdfd):
    """notebooks/Example_simulations.ipynb Example 3 (1500x200, 1 sideband) solved on the GPU; the flux consumer
    (flux.jl:37-47) must reproduce the notebook's printed outputs (cells 21, 23, 25)."""
    w, Om = 2 * math.pi * 1.939e14, 4.541e14
    a, q = 0.2202, 2.9263
    g = fdfd.Grid(0.01, [15, 10], [0.0, 15.0], [-1.0, 1.0])
    d = fdfd.ModulatedDevice(g, w, Om, 1)
    fdfd.setup_eps_r(d, lambda x, y: -a / 2 <= y <= a / 2, 12.25)
    fdfd.setup_deps_r(d, lambda x, y: (1.5 <= x <= 11.7) and (-a / 2 <= y <= 0), lambda x, y: np.exp(1j * q * x))
    fdfd.add_mode(d, fdfd.Mode(fdfd.TM, fdfd.XHAT, 3.5, fdfd.Point(0.2, 0), 4 * a))
    f = fdfd.solve(d)[0]
    P = fdfd.Point
    nin = fdfd.flux_surface_integral(f[1], P(1.25, 0), np.inf, fdfd.XHAT) / w
    nout = (fdfd.flux_surface_integral(f[2], P(11.95, 0), np.inf, fdfd.XHAT) / (Om + w)
            + fdfd.flux_surface_integral(f[1], P(11.95, 0), np.inf, fdfd.XHAT) / w
            + fdfd.flux_surface_integral(f[0], P(11.95, 0), np.inf, fdfd.XHAT) / (w - Om))
    assert abs(nin / 1.3625216010889075e-20 - 1) < 1e-6
    assert abs(nout / 1.3618731014650896e-20 - 1) < 1e-6
    assert abs(nout / nin - 0.9995240445191477) < 1e-6


def _ring(fdfd, dh):
    gargs = (dh, [15, 15], [-2.0, 2.0], [-2.0, 2.0])
    g, go = fdfd.Grid(*gargs), O.Grid2D(*gargs)
    w = 2 * math.pi * 200e12
    d = fdfd.Device(g, w)
    fdfd.setup_eps_r(d, [fdfd.Cylinder((0, 0), 0.8, 1.0), fdfd.Cylinder((0, 0), 1.0, 12.25)])  # notebook cell 31
    do = O.Device(go, [w])
    do.eps_r[:] = d.eps_r
    return d, do, g, go


def _match(got, ref):
    """pair each computed eigenvalue with its nearest reference (degenerate pairs allowed), return max rel error"""
    ref = list(ref)
    worst = 0.0
    for z in got:
        k = int(np.argmin([abs(z - r) for r in ref]))
        worst = max(worst, abs(z - ref[k]) / abs(ref[k]))
        ref.pop(k)
    return worst


@pytest.mark.parametrize("pol", ["TM", "TE"])
def test_eigenfrequency_ring(fdfd, pol):
    d, do, g, go = _ring(fdfd, 0.02)  # 200 x 200
    P = fdfd.TM if pol == "TM" else fdfd.TE
    Po = O.TM if pol == "TM" else O.TE
    nev = 4
    om, fields = fdfd.eigenfrequency(d, P, nev, which="LM")
    omo, fo = O.eigenfrequency(do, Po, nev + 2, which="LM", v0=np.ones(len(go), dtype=complex))
    assert _match(om, omo) <= EIG_TOL
    # eigen-pair residual with the oracle's matrix: ||A x - lambda x|| / ||lambda x||
    A, sigma, aux = O.eigen_matrix(do, Po)
    eps0, mu0, _ = O.normalize_parameters(go)
    for i in range(nev):
        x = fields[i].data[:, :, 0].ravel(order="F")
        lam = -(om[i] ** 2) * mu0 * (eps0 if pol == "TM" else 1.0)
        assert np.linalg.norm(A @ x - lam * x) / np.linalg.norm(lam * x) < 1e-6
        assert abs(np.linalg.norm(x) - 1) < 1e-8  # unit-norm vectors like ARPACK
    # recovered H/E components follow eigen.jl:90-91 / 108-109 applied to the same eigenvector
    _, fref = O.eigen_fields(do, Po, np.array([-(om[0] ** 2) * mu0 * (eps0 if pol == "TM" else 1.0)]),
                             fields[0].data[:, :, 0].ravel(order="F")[:, None], aux)
    assert rel(fields[0].data, fref[0]["data"]) < 1e-10


@pytest.mark.parametrize("pol", ["TM", "TE"])
def test_eigenfrequency_ring_matches_committed_fixture(fdfd, pol):
    """same device and call as test_eigenfrequency_ring, compared with tests/golden/eig_ring.json (oracle values committed with
    their generator, carrying the SURVEY 8c regression values at 400^2) instead of a live oracle run"""
    import json
    import os
    z = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eig_ring.json")))
    ref = [complex(a, b) for a, b in z[f"{pol}_200"]]
    d, _, _, _ = _ring(fdfd, 0.02)
    om, _ = fdfd.eigenfrequency(d, fdfd.TM if pol == "TM" else fdfd.TE, 4, which="LM")
    assert _match(om, ref) <= EIG_TOL


def _config4():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eig_config4.json")))


@pytest.mark.parametrize("pol", ["TM", "TE"])
def test_config4_ring_400_ten_modes(fdfd, pol):
    """BASELINE config 4 as stated: eigenfrequency() of the notebook ring (cell 31) at 400 x 400, the 10 modes nearest 200 THz
    (`which = :LM` on the shift-inverted spectrum, eigen.jl:86,104), against the committed oracle values (tests/golden/eig_config4.json,
    12 stored so that a degenerate pair at the cut cannot mismatch).  ncv is Arpack's default max(20, 2 nev + 1) = 21: ten wanted pairs
    (whispering-gallery modes with Q up to 2e6 next to Q ~ 7 leaky ones, two degenerate pairs) do not converge in 21 Arnoldi steps, so
    the Krylov-Schur restarts of csrc/arnoldi.cu are exercised and the basis stays at 22 vectors."""
    ref = [complex(a, b) for a, b in _config4()[f"ring_{pol}_400"]]
    d, _, _, _ = _ring(fdfd, 0.01)
    assert d.grid.N == (400, 400)
    om, fields = fdfd.eigenfrequency(d, fdfd.TM if pol == "TM" else fdfd.TE, 10, which="LM")
    assert len(om) == 10 and _match(om, ref) <= EIG_TOL
    assert fields[0].info["restarts"] > 21          # Arnoldi steps (operator applications): more than one basis' worth
    for f in fields:
        assert abs(np.linalg.norm(f.data[:, :, 0]) - 1) < 1e-8


def test_config4_photonic_crystal_cavity_ten_modes(fdfd):
    """BASELINE config 4, second device: an L3-type cavity (three rods removed from an 11 x 9 square lattice of eps = 12.25 rods,
    a = 0.5 um, r = 0.1 um; 256 x 224 grid), TM, the 10 modes nearest 200 THz -- the cavity mode at 191 THz (Q 1e4) and band-edge modes"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)           # the generator of the fixture builds the permittivity map for the test too
    phc_cavity_eps = mg.phc_cavity_eps
    ref = [complex(a, b) for a, b in _config4()["phc_TM"]]
    gargs = (0.025, [15, 15], [-3.2, 3.2], [-2.8, 2.8])
    g, go = fdfd.Grid(*gargs), O.Grid2D(*gargs)
    d = fdfd.Device(g, 2 * math.pi * 200e12)
    d.eps_r = phc_cavity_eps(go)
    assert g.N == (256, 224)
    om, _ = fdfd.eigenfrequency(d, fdfd.TM, 10, which="LM")
    assert _match(om, ref) <= EIG_TOL
