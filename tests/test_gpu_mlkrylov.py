"""GPU tests of the multilevel Krylov solver (csrc/mlkrylov.cu, FDFD_SOLVER_MLKRYLOV).

First run on hardware in round 2 (17/17 green); its arithmetic cores -- least-squares solve, grid transfers -- are also
checked on the CPU by tests/test_cabi_cpu.py.  Bars are those of the default solver: true relative residual
<= 1e-10 of the reference operator, fields within 1e-6 relative L2 of the oracle's direct solve."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W200 = 2 * math.pi * 200e12
FIELD_TOL = 1e-6
RES_TOL = 1e-10


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def spec(k1, k2=0, k3=0, restart=0):
    return k1 | (k2 << 8) | (k3 << 16) | (restart << 24)


def waveguide(fdfd, Nx, Ny, npml=(15, 10), dh=0.02):
    g = fdfd.Grid(dh, list(npml), [0.0, Nx * dh], [-Ny * dh / 2, Ny * dh / 2])
    d = fdfd.Device(g, W200)
    fdfd.setup_eps_r(d, lambda x, y: abs(y) <= 0.15, 12.0)
    fdfd.setup_src(d, fdfd.Point(0.6, 0.0), fdfd.XHAT)
    return d


def oracle_fields(d):
    from oracle import fdfd_oracle as O
    g = d.grid
    go = O.Grid2D(0.02, list(g.Npml), [g.bounds[0][0], g.bounds[1][0]], [g.bounds[0][1], g.bounds[1][1]])
    do = O.Device(go, [d.omega[0]])
    do.eps_r[:] = d.eps_r
    do.src[:] = d.src
    return O.solve(do, O.TM)["data"]


@pytest.mark.parametrize("size", [(192, 128), (201, 101), (250, 100)])      # even, odd (the transfers' short last edge), mixed
@pytest.mark.parametrize("ml_spec", [0, spec(8), spec(4, 6, 6)])          # defaults (6,12), two levels, four levels
def test_mlkrylov_waveguide_vs_oracle(fdfd, size, ml_spec):
    d = waveguide(fdfd, *size)
    f = fdfd.solve(d, fdfd.TM, solver=fdfd._lib.SOLVER_MLKRYLOV, ml_spec=ml_spec)
    assert f.info["flag"] == 0 and f.info["relres"] <= RES_TOL
    assert rel(f.data, oracle_fields(d)) <= FIELD_TOL


def test_mlkrylov_vs_default_solver_synthetic(fdfd):
    """512^2 synthetic map (bench workload at reduced size): same field as BiCGSTAB + multigrid, fewer fine cycles"""
    from fdfd_jl_b200 import workloads
    d = workloads.synthetic_tm_device(fdfd, 512, 512, density=1.0 / 160.0)
    ref = fdfd.solve(d, fdfd.TM)
    f = fdfd.solve(d, fdfd.TM, solver=fdfd._lib.SOLVER_MLKRYLOV)
    assert f.info["flag"] == 0 and f.info["relres"] <= RES_TOL
    assert rel(f.data, ref.data) <= FIELD_TOL
    # CPU prototype (tools/multilevel_prototype.py 512 300,6,-12): ~30 outer iterations vs ~93 BiCGSTAB iterations
    assert f.info["iters"] <= ref.info["iters"]


def test_mlkrylov_graph_and_plain_launches_agree(fdfd):
    d = waveguide(fdfd, 256, 128)
    a = fdfd.solve(d, fdfd.TM, solver=fdfd._lib.SOLVER_MLKRYLOV, use_graph=1)
    b = fdfd.solve(d, fdfd.TM, solver=fdfd._lib.SOLVER_MLKRYLOV, use_graph=0)
    assert a.info["flag"] == 0 and b.info["flag"] == 0
    assert a.info["iters"] == b.info["iters"]
    assert rel(a.data, b.data) <= 1e-12


def test_mlkrylov_zero_source_and_bad_arguments(fdfd):
    g = fdfd.Grid(0.02, [10, 10], [0.0, 2.56], [0.0, 2.56])
    d = fdfd.Device(g, W200)
    f = fdfd.solve(d, fdfd.TM, solver=fdfd._lib.SOLVER_MLKRYLOV)
    assert f.info["flag"] == 0 and np.all(f.data == 0)
    with pytest.raises(fdfd.FdfdError):
        fdfd.solve(d, fdfd.TM, solver=fdfd._lib.SOLVER_MLKRYLOV, mg_precision=1)     # fp32 multigrid only


def test_mlkrylov_te_vs_oracle(fdfd):
    """TE (driven.jl:45-51): the level operators carry the inverse averaged eps of their level"""
    from oracle import fdfd_oracle as O
    g = fdfd.Grid(0.02, [15, 15], [0.0, 3.84], [0.0, 2.56])
    d = fdfd.Device(g, W200)
    fdfd.setup_eps_r(d, [fdfd.Cylinder((1.9, 1.3), 0.5, 6.0)])
    fdfd.setup_src(d, fdfd.Point(0.8, 1.0))
    f = fdfd.solve(d, fdfd.TE, solver=fdfd._lib.SOLVER_MLKRYLOV)
    assert f.info["flag"] == 0 and f.info["relres"] <= RES_TOL
    go = O.Grid2D(0.02, [15, 15], [0.0, 3.84], [0.0, 2.56])
    do = O.Device(go, [W200]); do.eps_r[:] = d.eps_r; do.src[:] = d.src
    assert rel(f.data, O.solve(do, O.TE)["data"]) <= FIELD_TOL


def test_mlkrylov_device_cores_match_host(fdfd, ctx):
    """the device kernels of the least-squares solve and of the grid transfers against their host twins (which
    tests/test_cabi_cpu.py checks against NumPy): first thing to look at if a multilevel solve misbehaves"""
    import ctypes as C
    from fdfd_jl_b200._lib import ptr
    L = fdfd.lib()
    rng = np.random.default_rng(5)
    for k in (1, 6, 12, 40):
        H = np.zeros((k + 1, k), complex)
        for j in range(k):
            H[:j + 2, j] = rng.standard_normal(j + 2) + 1j * rng.standard_normal(j + 2)
        Hf = np.asfortranarray(H)
        yh, yd = np.zeros(k, complex), np.zeros(k, complex)
        rh, rd = C.c_double(), C.c_double()
        assert L.fdfd_debug_ml_lsq(k, ptr(Hf), 0.8, ptr(yh), C.byref(rh)) == 0
        fdfd._lib.check(L.fdfd_debug_ml_lsq_gpu(ctx.handle, k, ptr(Hf), 0.8, ptr(yd), C.byref(rd)), ctx.handle)
        assert np.abs(yh - yd).max() <= 1e-9 * max(1.0, np.abs(yh).max()) and abs(rh.value - rd.value) <= 1e-12   # fma contraction differs
    for nx, ny in ((64, 48), (65, 33), (256, 255)):
        ncx, ncy = (nx + 1) // 2, (ny + 1) // 2
        v = np.asfortranarray(rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny)))
        y = np.asfortranarray(rng.standard_normal((ncx, ncy)) + 1j * rng.standard_normal((ncx, ncy)))
        for mode, src, shape in ((0, v, (ncx, ncy)), (1, y, (nx, ny))):
            oh, od = np.zeros(shape, complex, order="F"), np.zeros(shape, complex, order="F")
            assert L.fdfd_debug_ml_transfer(nx, ny, mode, 0.25, ptr(src), ptr(oh)) == 0
            fdfd._lib.check(L.fdfd_debug_ml_transfer_gpu(ctx.handle, nx, ny, mode, 0.25, ptr(src), ptr(od)), ctx.handle)
            assert np.abs(oh - od).max() <= 1e-15


def test_mlkrylov_fused_and_modified_gram_schmidt_agree(fdfd, monkeypatch):
    """inner levels: the fused classical Gram-Schmidt pass (default) against modified Gram-Schmidt (FDFD_ML_MGS=1)"""
    d = waveguide(fdfd, 256, 128)
    a = fdfd.solve(d, fdfd.TM, solver=fdfd._lib.SOLVER_MLKRYLOV)
    monkeypatch.setenv("FDFD_ML_MGS", "1")
    b = fdfd.solve(d, fdfd.TM, solver=fdfd._lib.SOLVER_MLKRYLOV)
    assert a.info["flag"] == 0 and b.info["flag"] == 0
    assert abs(a.info["iters"] - b.info["iters"]) <= 2
    assert rel(a.data, b.data) <= FIELD_TOL


def test_eigenfrequency_with_multilevel_inner_solves(fdfd):
    """shift-invert Arnoldi (eigen.jl:69-96) whose inner solves run the multilevel Krylov solver (tol 1e-11)"""
    g = fdfd.Grid(0.02, [15, 15], [-2.0, 2.0], [-2.0, 2.0])
    d = fdfd.Device(g, W200)
    fdfd.setup_eps_r(d, [fdfd.Cylinder((0, 0), 0.8, 1.0), fdfd.Cylinder((0, 0), 1.0, 12.25)])  # notebook cell 31
    om_ref, _ = fdfd.eigenfrequency(d, fdfd.TM, 6, which="LM")
    om, _ = fdfd.eigenfrequency(d, fdfd.TM, 4, which="LM", solver=fdfd._lib.SOLVER_MLKRYLOV)
    ref = list(om_ref)
    for z in om:
        k = int(np.argmin([abs(z - r) for r in ref]))
        assert abs(z - ref[k]) / abs(ref[k]) <= 1e-8
        ref.pop(k)


def test_mlkrylov_thin_pml(fdfd):
    """dh = 0.01, Npml = 10: the PML is thinner than a level-2 cell, the case that needed Galerkin-consistent multigrid
    transfers.  The deflation transfers are plain bilinear; flexible GMRES must still converge."""
    g = fdfd.Grid(0.01, [10, 10], [0.0, 3.84], [-0.96, 0.96])
    d = fdfd.Device(g, W200)
    fdfd.setup_eps_r(d, lambda x, y: abs(y) <= 0.15, 12.0)
    fdfd.setup_src(d, fdfd.Point(0.6, 0.0), fdfd.XHAT)
    ref = fdfd.solve(d, fdfd.TM)
    f = fdfd.solve(d, fdfd.TM, solver=fdfd._lib.SOLVER_MLKRYLOV)
    assert ref.info["flag"] == 0 and f.info["flag"] == 0 and f.info["relres"] <= RES_TOL
    assert rel(f.data, ref.data) <= FIELD_TOL
