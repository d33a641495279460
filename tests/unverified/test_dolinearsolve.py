"""GPU tests of csrc/linsolve.cu: fdfd_dolinearsolve_csc, the dolinearsolve(A, b, matrixsym) seam of the reference
(src/solver/solver.jl:4-41) for callers that hand over an assembled SparseMatrixCSC.

NOT part of `-m gpu`: written in a session without GPU access, compiled only (the host half -- CSC -> SELL-32 -- is covered by
tests/test_cabi_cpu.py).  Run with  FDFD_RUN_UNVERIFIED=1 python -m pytest tests/unverified/test_dolinearsolve.py -x -q --timeout 900
on a B200; once green, move into tests/test_gpu_parity.py with the `gpu` marker.
Bars: true relative residual <= 1e-10 (recomputed here with SciPy), solution / fields within 1e-6 relative L2 of the oracle's
sparse direct solve."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import fdfd_oracle as O

pytestmark = [pytest.mark.gpu_unverified,
              pytest.mark.skipif(os.environ.get("FDFD_RUN_UNVERIFIED") != "1", reason="unverified GPU path: set FDFD_RUN_UNVERIFIED=1 on a GPU box")]

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
