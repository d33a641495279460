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
