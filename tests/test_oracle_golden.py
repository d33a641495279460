"""Pins the CPU oracle to the reference's only known-answer numbers: the photon-number outputs of
notebooks/Example_simulations.ipynb cells 21/23/25 (Example 3, modulated TM waveguide, 1500x200x3 unknowns).
~100 s on 8 cores (one SuperLU factorisation of 9e5 unknowns)."""
import math

import numpy as np
import pytest

from oracle import fdfd_oracle as O

NIN_REF = 1.3625216010889075e-20   # cell 21
NOUT_REF = 1.3618731014650896e-20  # cell 23
RATIO_REF = 0.9995240445191477     # cell 25


def build_example3():
    g = O.Grid2D(0.01, [15, 10], [0.0, 15.0], [-1.0, 1.0])
    w, Om = 2 * math.pi * 1.939e14, 4.541e14
    d = O.ModulatedDevice(g, [w], Omega=Om, nsidebands=1)
    a, q = 0.2202, 2.9263
    O.mask_values(d.eps_r, g, lambda x, y: -a / 2 <= y <= a / 2, 12.25)
    O.mask_values(d.deps_r, g, lambda x, y: (1.5 <= x <= 11.7) and (-a / 2 <= y <= 0), lambda x, y: np.exp(1j * q * x))
    d.modes.append(O.Mode(O.TM, O.X, 3.5, (0.2, 0), 4 * a))
    return g, d, w, Om


@pytest.mark.slow
def test_notebook_photon_numbers():
    g, d, w, Om = build_example3()
    assert g.N == (1500, 200)
    f = O.solve_modulated(d)[0]
    nin = O.flux_surface_integral_tm_x(g, f[1]["data"], (1.25, 0), np.inf) / w
    nout = (O.flux_surface_integral_tm_x(g, f[2]["data"], (11.95, 0), np.inf) / (Om + w)
            + O.flux_surface_integral_tm_x(g, f[1]["data"], (11.95, 0), np.inf) / w
            + O.flux_surface_integral_tm_x(g, f[0]["data"], (11.95, 0), np.inf) / (w - Om))
    assert abs(nin / NIN_REF - 1) < 1e-10
    assert abs(nout / NOUT_REF - 1) < 1e-10
    assert abs(nout / nin - RATIO_REF) < 1e-10


def test_mode_slice_intermediates():
    """SURVEY §4 intermediate values of the same example: slice indices and the 1-D mode's beta."""
    g, d, w, Om = build_example3()
    beta, vec, ix, iy = O.get_modes(d, O.TM, w, 3.5, 1, (0.2, 0), O.X, 4 * 0.2202)
    assert ix + 1 == 21 and iy[0] + 1 == 57 and iy[-1] + 1 == 145 and len(iy) == 89
    assert abs(beta[0].real - 11.54702336) < 1e-6


def test_operator_identities():
    """closed-form 5-point coefficients == the assembled sparse products, f.b and b.f orderings."""
    g = O.Grid2D(0.05, [6, 5], [0, 2.0], [0, 1.5])
    w = 2 * math.pi * 200e12
    d = O.Device(g, [w])
    rng = np.random.default_rng(0)
    d.eps_r = (1 + 11 * rng.random(g.size())) + 0j
    A, b, _ = O.system_matrix(d, w, O.TM)
    eps0, mu0, _ = O.normalize_parameters(g)
    cxm, cxp, cym, cyp = O.stencil_coefficients(g, w, "fb")
    Nx, Ny = g.size()
    x = rng.standard_normal((Nx, Ny)) + 1j * rng.standard_normal((Nx, Ny))
    y = (cxm[:, None] * (np.roll(x, 1, 0) - x) + cxp[:, None] * (np.roll(x, -1, 0) - x)
         + cym[None, :] * (np.roll(x, 1, 1) - x) + cyp[None, :] * (np.roll(x, -1, 1) - x) + w ** 2 * eps0 * d.eps_r * x)
    ref = (A @ x.ravel(order="F")).reshape((Nx, Ny), order="F")
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-13
    # A is not complex-symmetric, diag(sxf*syf) A is
    sxf, sxb, syf, syb = O.inv_sfactors(g, w)
    Dm = np.outer(1 / sxf, 1 / syf).ravel(order="F")
    import scipy.sparse as sp
    As = sp.diags(Dm) @ A
    assert abs(As - As.T).max() / abs(As).max() < 1e-14
    assert abs(A - A.T).max() / abs(A).max() > 1e-3
