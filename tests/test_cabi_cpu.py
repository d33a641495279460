"""CPU-side checks of the drop-in boundary: the shared library builds, loads, exports every symbol that
include/fdfd_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(fdfd):
    L = fdfd.lib()
    hdr = open(os.path.join(ROOT, "include", "fdfd_b200.h")).read()
    declared = set(re.findall(r"\b(fdfd_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "header parse failed"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/fdfd_b200.h but not exported"
    assert declared == set(fdfd._lib.EXPORTS)
    assert L.fdfd_abi_version() == 1


def test_struct_layouts_match_header(fdfd):
    assert ctypes.sizeof(fdfd.GridT) == 4 * 8 + 5 * 8
    assert ctypes.sizeof(fdfd.SolveOpts) == 104
    assert ctypes.sizeof(fdfd.Info) == 56
    o = fdfd.default_opts()
    assert o.tol == 1e-10 and o.precond == fdfd._lib.PRECOND_MG and o.mg_beta == 0.5


def test_no_cpu_fallback(fdfd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fdfd.FdfdError) as e:
        fdfd.Context(0)
    assert "no CPU path" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fdfd.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "fdfd_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_host_mirror_grid_matches_oracle(fdfd):
    from oracle import fdfd_oracle as O
    import numpy as np
    for args in [(0.02, [15, 10], [0.0, 10.0], [-1.0, 1.0]), (0.01, [15, 15], [-3, 3], [-3, 3]), (0.03, [4, 7], [0, 1.0], [0, 2.0])]:
        g, go = fdfd.Grid(*args), O.Grid2D(*args)
        assert g.N == go.N and g.Npml == go.Npml
        assert np.array_equal(fdfd.xc(g), O.xc(go)) and np.array_equal(fdfd.yc(g), O.yc(go))
        for p in [(0.0, 0.0), (1.0, 0.3), (-5, 7), (0.995, -0.505)]:
            assert fdfd.coord2ind(g, p) == O.coord2ind(go, p)
    # mode source (host side in both): identical slice and vector
    g = fdfd.Grid(0.02, [15, 10], [0.0, 10.0], [-1.0, 1.0]); go = O.Grid2D(0.02, [15, 10], [0.0, 10.0], [-1.0, 1.0])
    w = 2 * np.pi * 200e12
    d, do = fdfd.Device(g, w), O.Device(go, [w])
    fdfd.setup_eps_r(d, [fdfd.Box((5.0, 0.0), (np.inf, 0.3), 12)])
    O.compose_shapes(do.eps_r, go, [(O.box_region((5.0, 0.0), (np.inf, 0.3)), 12)])
    assert np.array_equal(d.eps_r, do.eps_r)
    fdfd.add_mode(d, fdfd.Mode(fdfd.TM, fdfd.XHAT, 3.5, fdfd.Point(1.0, 0), 0.8)); do.modes.append(O.Mode(O.TM, O.X, 3.5, (1.0, 0), 0.8))
    fdfd._apply_modes(d, w); O._apply_modes(do, w)
    assert np.allclose(d.src, do.src, rtol=1e-12, atol=1e-15) and abs(np.linalg.norm(d.src) - 1) < 1e-12


def test_hessenberg_eigensolver_host(fdfd):
    """the small dense solver behind fdfd_eigenfrequency's Ritz pairs (host code, no GPU)"""
    import numpy as np
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 40, 120):
        H = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)), -1)
        Hf = np.asfortranarray(H); ev = np.empty(n, complex); vec = np.empty((n, n), complex, order="F")
        assert fdfd.lib().fdfd_debug_hess_eig(n, fdfd.ptr(Hf), fdfd.ptr(ev), fdfd.ptr(vec)) == 0
        ref = np.linalg.eigvals(H)
        assert max(min(abs(e - ref)) for e in ev) < 1e-10 * max(1, abs(ref).max())
        for k in range(n):
            assert np.linalg.norm(H @ vec[:, k] - ev[k] * vec[:, k]) < 1e-10 * max(1, abs(ref).max())


def test_general_eig_host(fdfd):
    """eigen-solver of a general small matrix (Householder -> Hessenberg -> shifted QR): what the Ritz pairs come from after a thick restart"""
    import numpy as np
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 9, 41):
        A = np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        ev = np.empty(n, complex); vec = np.empty((n, n), complex, order="F")
        assert fdfd.lib().fdfd_debug_general_eig(n, fdfd.ptr(A), fdfd.ptr(ev), fdfd.ptr(vec)) == 0
        ref = np.linalg.eigvals(A)
        assert max(min(abs(e - ref)) for e in ev) < 1e-10 * max(1, abs(ref).max())
        for k in range(n):
            assert np.linalg.norm(A @ vec[:, k] - ev[k] * vec[:, k]) < 1e-9 * max(1, abs(ref).max())


def test_krylov_schur_host(fdfd):
    """the Krylov-Schur loop of fdfd_eigenfrequency (thick restarts keep the basis at ncv + 1 vectors like Arpack, eigen.jl:86) on dense
    operators with host vectors: non-normal spectrum that needs several restarts, a degenerate pair (the ring resonator's case), every
    `which`, and an operator of rank 3 whose Krylov space closes after 3 steps (invariant-subspace restart)"""
    import numpy as np
    L = fdfd.lib()
    rng = np.random.default_rng(5)

    def run(OP, nev, ncv, which, tol=1e-12, max_steps=2000):
        n = OP.shape[0]
        nu = np.empty(nev, complex); vec = np.empty((n, nev), complex, order="F")
        steps, restarts = ctypes.c_int32(), ctypes.c_int32()
        code = L.fdfd_debug_krylov_schur(n, fdfd.ptr(np.asfortranarray(OP)), nev, ncv, which, tol, max_steps, fdfd.ptr(nu), fdfd.ptr(vec),
                                         ctypes.byref(steps), ctypes.byref(restarts))
        return code, nu, vec, steps.value, restarts.value

    n = 300
    X = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    lam = np.concatenate([[5.0, 4.5 + 1j, 4.5 + 1j, -4.2j, 3.9, -3.7], 2.0 * (rng.random(n - 6) - 0.5) + 1.5j * (rng.random(n - 6) - 0.5)])
    OP = X @ np.diag(lam) @ np.linalg.inv(X)
    code, nu, vec, steps, restarts = run(OP, 6, 14, 0)
    assert code == 0 and restarts >= 1                       # 14 vectors are not enough without restarting
    assert np.allclose(sorted(abs(nu))[::-1], sorted(abs(lam[:6]))[::-1], rtol=1e-9)
    for k in range(6):
        v = vec[:, k]
        assert np.linalg.norm(OP @ v - nu[k] * v) <= 1e-8 * abs(nu[k]) * np.linalg.norm(v)
    # every `which` picks the matching end of the spectrum
    keys = {0: abs(lam), 1: lam.real, 2: -lam.real, 3: lam.imag, 4: -lam.imag}
    for which, key in keys.items():
        code, nu, _, _, _ = run(OP, 2, 20, which, tol=1e-10)
        want = lam[np.argsort(-key)[:2]]
        assert code == 0
        for z in nu:
            assert min(abs(z - want)) <= 1e-7 * max(1.0, abs(z)), (which, nu, want)
    # rank-3 operator: the Krylov space is exhausted after 3-4 steps; the two nonzero-eigenvalue pairs asked for are exact
    U = rng.standard_normal((40, 3)) + 1j * rng.standard_normal((40, 3))
    Wt = rng.standard_normal((3, 40)) + 1j * rng.standard_normal((3, 40))
    OP3 = U @ Wt
    code, nu, _, steps, _ = run(OP3, 2, 12, 0, tol=1e-10)
    ref = np.linalg.eigvals(Wt @ U)
    ref = ref[np.argsort(-abs(ref))]
    assert code == 0 and steps <= 13
    for z, r in zip(sorted(nu, key=lambda t: -abs(t)), ref[:2]):
        assert abs(z - r) <= 1e-8 * abs(r)
    # an unreachable tolerance ends in FDFD_ERR_NOCONV, not in an unbounded basis
    code, *_ = run(OP + 1e-3 * rng.standard_normal((n, n)), 6, 9, 0, tol=1e-30, max_steps=60)
    assert code == 3


def test_mlkrylov_least_squares_core_host(fdfd):
    """the one-thread Givens least-squares solve of the multilevel Krylov solver (csrc/mlkrylov.cu), run on the host"""
    import numpy as np
    from fdfd_jl_b200._lib import ptr
    L = fdfd.lib()
    rng = np.random.default_rng(3)
    for k in (1, 2, 7, 40, 96):
        H = np.zeros((k + 1, k), complex)
        for j in range(k):
            H[:j + 2, j] = rng.standard_normal(j + 2) + 1j * rng.standard_normal(j + 2)
        Hf = np.asfortranarray(H)
        y = np.zeros(k, complex)
        res = ctypes.c_double()
        assert L.fdfd_debug_ml_lsq(k, ptr(Hf), 1.75, ptr(y), ctypes.byref(res)) == 0
        g = np.zeros(k + 1, complex)
        g[0] = 1.75
        yr = np.linalg.lstsq(H, g, rcond=None)[0]
        assert np.abs(y - yr).max() <= 1e-13 * np.linalg.cond(H) * max(1.0, np.abs(yr).max())   # random Hessenberg matrices are ill-conditioned for large k
        assert abs(np.linalg.norm(g - H @ y) - np.linalg.norm(g - H @ yr)) <= 1e-12
        assert abs(res.value - np.linalg.norm(g - H @ yr)) <= 1e-12
    # a zero column (lucky breakdown) must not produce NaNs
    H = np.asfortranarray(np.zeros((3, 2), complex)); H[0, 0] = 2.0
    y = np.zeros(2, complex)
    assert L.fdfd_debug_ml_lsq(2, ptr(H), 1.0, ptr(y), None) == 0 and np.all(np.isfinite(y)) and abs(y[0] - 0.5) < 1e-15


def test_mlkrylov_transfers_host(fdfd):
    """Z (bilinear, coarse I <-> fine 2I, periodic) and Z^T of the multilevel Krylov solver are adjoint, Z reproduces
    constants, and Z^T Z / 4 has unit row sums on even grids -- checked on the host for even, odd and mixed sizes"""
    import numpy as np
    from fdfd_jl_b200._lib import ptr
    L = fdfd.lib()
    rng = np.random.default_rng(4)
    for nx, ny in ((8, 6), (9, 7), (8, 7), (5, 4), (64, 33)):
        ncx, ncy = (nx + 1) // 2, (ny + 1) // 2
        v = np.asfortranarray(rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny)))
        y = np.asfortranarray(rng.standard_normal((ncx, ncy)) + 1j * rng.standard_normal((ncx, ncy)))
        zt = np.zeros((ncx, ncy), complex, order="F")
        z = np.zeros((nx, ny), complex, order="F")
        assert L.fdfd_debug_ml_transfer(nx, ny, 0, 1.0, ptr(v), ptr(zt)) == 0
        assert L.fdfd_debug_ml_transfer(nx, ny, 1, 1.0, ptr(y), ptr(z)) == 0
        assert abs(np.vdot(z, v) - np.vdot(y, zt)) <= 1e-12 * np.linalg.norm(z) * np.linalg.norm(v)
        one = np.asfortranarray(np.ones((ncx, ncy), complex))
        assert L.fdfd_debug_ml_transfer(nx, ny, 1, 1.0, ptr(one), ptr(z)) == 0 and np.abs(z - 1).max() == 0
        if nx % 2 == 0 and ny % 2 == 0:
            ones_f = np.asfortranarray(np.ones((nx, ny), complex))
            assert L.fdfd_debug_ml_transfer(nx, ny, 0, 0.25, ptr(ones_f), ptr(zt)) == 0 and np.abs(zt - 1).max() < 1e-15


def test_dolinearsolve_sell_transposition_host(fdfd):
    """fdfd_dolinearsolve_csc's host half (solver.jl:4 seam): the CSC -> SELL-32 transposition and the per-row summation the SpMV
    kernel shares with this hook reproduce A @ x and 1/diag(A) for ragged, empty-row, duplicate-entry and n % 32 != 0 matrices in
    both index bases, and for the reference's own TM system matrix (oracle assembly, driven.jl:35); malformed CSC is refused."""
    import numpy as np
    import scipy.sparse as sp
    from oracle import fdfd_oracle as O
    rng = np.random.default_rng(7)
    for n, dens in ((1, 1.0), (5, 0.5), (32, 0.2), (33, 0.1), (257, 0.02), (1000, 0.006)):
        A = (sp.random(n, n, density=dens, random_state=rng) + 1j * sp.random(n, n, density=dens, random_state=rng)).tocsc()
        if n > 5:   # a diagonal with some exact zeros, and rows left empty by the random pattern
            A = (A + sp.diags((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (rng.random(n) > 0.3))).tocsc()
        A.sort_indices()
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        d = A.diagonal()
        dref = np.where(d != 0, 1 / np.where(d != 0, d, 1), 1)
        for base in (0, 1):
            y, dinv, pad = fdfd._sell_spmv_host((A.indptr + base, A.indices + base, A.data), x, index_base=base)
            assert np.abs(y - A @ x).max() <= 1e-13 * max(1.0, np.abs(A @ x).max())
            assert np.abs(dinv * np.where(d != 0, d, 1) - 1).max() <= 1e-12 and np.all(dinv[d == 0] == 1)
            assert pad % 32 == 0 and pad >= A.nnz
            assert np.allclose(dinv, dref, rtol=1e-12)
    # duplicates are summed (Julia's sparse(I, J, V) semantics)
    colptr, rowval = np.array([0, 2, 3]), np.array([0, 0, 1])
    y, dinv, _ = fdfd._sell_spmv_host((colptr, rowval, np.array([1 + 1j, 2.0, 4.0])), np.array([1.0, 1.0]))
    assert np.allclose(y, [3 + 1j, 4.0]) and np.allclose(dinv, [1 / (3 + 1j), 0.25])
    # the reference's TM matrix (5 nnz/row, periodic wrap entries): SELL padding is zero when every row has the same length
    g = O.Grid2D(0.05, [6, 5], [0.0, 2.0], [-0.8, 0.8])
    dev = O.Device(g, [2 * np.pi * 200e12])
    dev.eps_r[10:20, 12:18] = 12.0
    A = O.system_matrix(dev, dev.omega[0], O.TM)[0].tocsc()
    A.sort_indices()
    x = rng.standard_normal(A.shape[0]) + 1j * rng.standard_normal(A.shape[0])
    y, _, pad = fdfd._sell_spmv_host(A, x)
    assert np.abs(y - A @ x).max() <= 1e-13 * np.abs(A @ x).max() and pad == 5 * A.shape[0]
    # grid-hinted path (fdfd_dolinearsolve_csc_grid): the permittivity is read back off the row sums, A 1 = w^2 eps0 L0 eps_r,
    # also after a Born step A + Diagonal(coeff |ez|^2) (nonlinear.jl:97) has changed the diagonal
    eps0 = O.normalize_parameters(g)[0]
    dev.eps_r[3, 4] = 2.5 + 0.1j
    A = O.system_matrix(dev, dev.omega[0], O.TM)[0].tocsc()
    rs = np.empty(A.shape[0], complex)
    fdfd._sell_spmv_host(A, x, rowsum=rs)
    assert np.abs(rs / (dev.omega[0] ** 2 * eps0) - dev.eps_r.ravel(order="F")).max() <= 1e-12
    kerr = 1e-3 * rng.random(A.shape[0])
    fdfd._sell_spmv_host((A + sp.diags(dev.omega[0] ** 2 * eps0 * kerr)).tocsc(), x, rowsum=rs)
    assert np.abs(rs / (dev.omega[0] ** 2 * eps0) - (dev.eps_r.ravel(order="F") + kerr)).max() <= 1e-12
    # ... and the operator rebuilt from (grid, omega, eps_eff) IS the caller's matrix, entry by entry: the premise of the fast path
    Ab = (A + sp.diags(dev.omega[0] ** 2 * eps0 * kerr)).tocsr()
    dev2 = O.Device(g, [dev.omega[0]])
    dev2.eps_r[:] = (rs / (dev.omega[0] ** 2 * eps0)).reshape(dev.eps_r.shape, order="F")
    A2 = O.system_matrix(dev2, dev.omega[0], O.TM)[0].tocsr()
    assert abs(Ab - A2).max() <= 1e-12 * abs(Ab).max()
    with pytest.raises(fdfd.FdfdError):
        fdfd._sell_spmv_host((np.array([0, 1, 2]), np.array([0, 5]), np.array([1.0, 1.0])), np.ones(2))
    with pytest.raises(fdfd.FdfdError):
        fdfd._sell_spmv_host((np.array([0, 2, 1]), np.array([0, 1]), np.array([1.0, 1.0])), np.ones(2))


def test_dolinearsolve_fails_loudly_without_gpu(fdfd):
    import numpy as np
    import scipy.sparse as sp
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fdfd.FdfdError) as e:
        fdfd.dolinearsolve(sp.identity(4, dtype=complex, format="csc"), np.ones(4))
    assert "no CPU path" in str(e.value)
    with pytest.raises(ValueError):
        fdfd.dolinearsolve(sp.random(3, 4, density=0.5, format="csc"), np.ones(4))


def test_batched_apply_host_checks_and_no_cpu_fallback(fdfd):
    """fdfd_apply_operator_batched: the host mirror checks the (Nx, Ny, B) layout; without a GPU the call fails loudly like every entry point"""
    import numpy as np
    import torch
    g = fdfd.Grid(0.1, [2, 2], [0, 1.0], [0, 0.8])
    eps = np.ones(g.N, dtype=complex)
    with pytest.raises(ValueError):
        fdfd.apply_operator_batched(g, fdfd.TM, 1e15, eps, np.ones(g.N, dtype=complex))          # 2-D: no batch axis
    with pytest.raises(ValueError):
        fdfd.apply_operator_batched(g, fdfd.TM, 1e15, eps, np.ones((3, 3, 2), dtype=complex))    # wrong grid
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fdfd.FdfdError) as e:
        fdfd.apply_operator_batched(g, fdfd.TM, 1e15, eps, np.ones(g.N + (2,), dtype=complex))
    assert "no CPU path" in str(e.value)
