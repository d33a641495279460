"""GPU tests of csrc/linsolve.cu: fdfd_dolinearsolve_csc, the dolinearsolve(A, b, matrixsym) seam of the reference
(src/solver/solver.jl:4-41) for callers that hand over an assembled SparseMatrixCSC.

First run on hardware in round 2 (7/7 green); the host half -- CSC -> SELL-32 -- is covered by tests/test_cabi_cpu.py.
Bars: true relative residual <= 1e-10 (recomputed here with SciPy), solution / fields within 1e-6 relative L2 of the oracle's
sparse direct solve."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import fdfd_oracle as O

pytestmark = pytest.mark.gpu

W200 = 2 * np.pi * 200e12
RES_TOL, FIELD_TOL = 1e-10, 1e-6


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def test_diagonally_dominant_random_matrix(fdfd):
    """ragged complex matrix, n not a multiple of 32, both index bases, device result == SciPy direct solve"""
    rng = np.random.default_rng(11)
    n = 1000
    A = (sp.random(n, n, density=0.01, random_state=rng) + 1j * sp.random(n, n, density=0.01, random_state=rng)).tocsc()
    A = (A + sp.diags(4.0 + rng.random(n) + 1j * rng.standard_normal(n))).tocsc()
    A.sort_indices()
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    ref = spla.spsolve(A, b)
    for base in (0, 1):
        x, info = fdfd.dolinearsolve((A.indptr + base, A.indices + base, A.data), b, index_base=base, return_info=True)
        assert info["flag"] == 0 and info["relres"] <= RES_TOL
        assert np.linalg.norm(b - A @ x) / np.linalg.norm(b) <= 2 * RES_TOL
        assert rel(x, ref) <= 1e-8
    x = fdfd.dolinearsolve(A, b, use_graph=0)   # eager launches give the same answer as the captured iteration
    assert rel(x, ref) <= 1e-8


def test_reference_driver_with_b200_linear_solver(fdfd):
    """the FDFD_SOLVER=b200 use of the seam: the reference's own driver (oracle restatement of driven.jl:4-59) assembles A and b and
    calls dolinearsolve; with the GPU solver plugged in the fields equal those of the sparse direct solve.  Small vacuum dipole
    (the device of test_jacobi_bicgstab_option: Jacobi-BiCGSTAB needs thousands of iterations on a Helmholtz matrix)."""
    go = O.Grid2D(0.06, [10, 10], [-3, 3], [-3, 3])   # 100 x 100
    do = O.Device(go, [W200])
    O.setup_src_point(do, (0, 0))
    infos = []

    def gpu_linsolve(A, b):
        x, info = fdfd.dolinearsolve(A, b, maxit=100000, check_every=64, return_info=True)
        infos.append(info)
        assert np.linalg.norm(b - A @ x) / np.linalg.norm(b) <= 2 * RES_TOL
        return x

    f = O.solve(do, O.TM, linsolve=gpu_linsolve)
    fo = O.solve(do, O.TM)
    assert infos and infos[0]["flag"] == 0 and infos[0]["relres"] <= RES_TOL
    assert rel(f["data"], fo["data"]) <= FIELD_TOL


def test_matches_the_matrix_free_jacobi_solve(fdfd):
    """the assembled-matrix path and the matrix-free path run the same BiCGSTAB + Jacobi: same solution on the library's own CSC
    assembly (1-based, as Julia would hand it over)"""
    g = fdfd.Grid(0.06, [10, 10], [-3, 3], [-3, 3])
    d = fdfd.Device(g, W200)
    fdfd.setup_src(d, fdfd.Point(0, 0))
    colptr, rowval, nzval = fdfd.assemble_system(g, fdfd.TM, W200, d.eps_r, fmt=fdfd._lib.CSC, index_base=1)
    b = 1j * W200 * np.asarray(d.src).ravel(order="F")
    x, info = fdfd.dolinearsolve((colptr, rowval, nzval), b, index_base=1, maxit=100000, check_every=64, return_info=True)
    f = fdfd.solve(d, fdfd.TM, precond=1, maxit=100000, check_every=64)
    assert info["flag"] == 0 and rel(x, f["Ez"].ravel(order="F")) <= FIELD_TOL


def test_bad_arguments_and_nonconvergence_are_reported(fdfd):
    A = sp.identity(8, dtype=complex, format="csc")
    with pytest.raises(fdfd.FdfdError):
        fdfd.dolinearsolve((np.array([0, 1, 2]), np.array([0, 7]), np.array([1.0, 1.0])), np.ones(2))
    with pytest.raises(ValueError):
        fdfd.dolinearsolve(A, np.ones(5))
    rng = np.random.default_rng(3)
    B = (sp.random(200, 200, density=0.05, random_state=rng) + sp.identity(200) * 1e-3).tocsc()   # far from diagonally dominant
    with pytest.raises(fdfd.FdfdError):
        fdfd.dolinearsolve(B, np.ones(200), maxit=5)
    assert np.allclose(fdfd.dolinearsolve(A, np.zeros(8)), 0)   # b = 0 -> x = 0


# ---- grid-hinted seam (fdfd_dolinearsolve_csc_grid): a matrix that IS the TM operator of the grid runs on the multigrid path -------
def _waveguide(fdfd):
    gargs = (0.02, [15, 10], [0.0, 5.0], [-1.0, 1.0])   # 250 x 100
    g, go = fdfd.Grid(*gargs), O.Grid2D(*gargs)
    do = O.Device(go, [W200])
    xs, ys = O.xc(go)[:, None], O.yc(go)[None, :]
    do.eps_r[(np.abs(ys) <= 0.15) & (xs >= 0)] = 12.0
    O.setup_src_point(do, (1.0, 0.0))
    return g, go, do


def test_grid_hint_runs_the_multigrid_path_on_the_reference_matrix(fdfd):
    g, go, do = _waveguide(fdfd)
    A, b, _ = O.system_matrix(do, W200, O.TM)
    A = A.tocsc()
    x, info = fdfd.dolinearsolve(A, b, grid=g, omega=W200, return_info=True)
    assert info["flag"] == 0 and info["mg_levels"] > 0 and info["relres"] <= RES_TOL
    assert np.linalg.norm(b - A @ x) / np.linalg.norm(b) <= 2 * RES_TOL
    assert rel(x, spla.spsolve(A, b)) <= FIELD_TOL
    # 1-based Julia arrays, and the reference driver on top of the seam
    f = O.solve(do, O.TM, linsolve=lambda A_, b_: fdfd.dolinearsolve(A_, b_, grid=g, omega=W200))
    assert rel(f["data"], O.solve(do, O.TM)["data"]) <= FIELD_TOL
    x1 = fdfd.dolinearsolve((A.indptr + 1, A.indices + 1, A.data), b, index_base=1, grid=g, omega=W200)
    assert rel(x1, x) <= 1e-8


def test_grid_hint_born_iteration_of_the_kerr_solver(fdfd):
    """_doborn (nonlinear.jl:92-106): ez <- dolinearsolve(A + Diagonal(coeff |ez|^2), b) until the step is small; every system is a TM
    operator with a modified permittivity, so every step takes the multigrid path; the iterates equal those of the direct solver"""
    g, go, do = _waveguide(fdfd)
    A, b, _ = O.system_matrix(do, W200, O.TM)
    A = A.tocsc()
    eps0 = O.normalize_parameters(go)[0]
    chi = np.where(do.eps_r.ravel(order="F").real > 1, 1.0, 0.0)
    ez_ref = spla.spsolve(A, b)
    coeff = W200 ** 2 * eps0 * 3 * chi * (0.05 / np.abs(ez_ref).max() ** 2)   # chi3 scaled to a 5 % index change at the field maximum
    ez, _ = fdfd.dolinearsolve(A, b, grid=g, omega=W200, return_info=True)
    for _ in range(3):
        An = (A + sp.diags(coeff * np.abs(ez_ref) ** 2)).tocsc()
        ez_ref = spla.spsolve(An, b)
        Ag = (A + sp.diags(coeff * np.abs(ez) ** 2)).tocsc()
        ez, info = fdfd.dolinearsolve(Ag, b, grid=g, omega=W200, return_info=True)
        assert info["flag"] == 0 and info["mg_levels"] > 0 and info["relres"] <= RES_TOL
        assert rel(ez, ez_ref) <= 1e-5


def test_grid_hint_other_ordering_and_fallback(fdfd):
    g = fdfd.Grid(0.06, [10, 10], [-3, 3], [-3, 3])
    d = fdfd.Device(g, W200)
    fdfd.setup_src(d, fdfd.Point(0, 0))
    b = 1j * W200 * np.asarray(d.src).ravel(order="F")
    # the b.f ordering of modulation.jl:82 is recognised too
    trip = fdfd.assemble_system(g, fdfd.TM, W200, d.eps_r, ordering=fdfd._lib.ORDER_BF, fmt=fdfd._lib.CSC, index_base=0)
    x, info = fdfd.dolinearsolve(trip, b, grid=g, omega=W200, return_info=True)
    assert info["flag"] == 0 and info["mg_levels"] > 0 and info["relres"] <= RES_TOL
    # a matrix that is NOT the operator of this grid (here 2 A: the row sums give 2 eps, the couplings do not match) falls back
    colptr, rowval, nzval = fdfd.assemble_system(g, fdfd.TM, W200, d.eps_r, fmt=fdfd._lib.CSC, index_base=0)
    x2, info2 = fdfd.dolinearsolve((colptr, rowval, 2 * nzval), b, grid=g, omega=W200, maxit=100000, check_every=64, return_info=True)
    assert info2["flag"] == 0 and info2["mg_levels"] == 0 and info2["relres"] <= RES_TOL
    xa = fdfd.dolinearsolve((colptr, rowval, nzval), b, grid=g, omega=W200)
    assert rel(2 * x2, xa) <= FIELD_TOL
