"""GPU tests of the row-slab sharded solve (csrc/slab.cu) through the C ABI.  All slabs live on ONE GPU here
(thread transport: one host thread, context and stream per slab, device-to-device halo copies), which exercises exactly
the slab layout, halo refreshes, lock-step multigrid and allreduced dot products that the NCCL transport runs with one
process per GPU (that transport is covered by tools/slab_nccl_check.py under torchrun on >= 2 GPUs).
Bars: true relative residual <= 1e-10 on every slab; fields within 1e-6 relative L2 of the oracle's direct solve and of
the single-GPU solve."""
import math

import numpy as np
import pytest

from oracle import fdfd_oracle as O

pytestmark = pytest.mark.gpu

W200 = 2 * math.pi * 200e12
FIELD_TOL = 1e-6
RES_TOL = 1e-10


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def waveguide(fdfd, Nx, Ny, npml=(15, 10), dh=0.02):
    g = fdfd.Grid(dh, list(npml), [0.0, Nx * dh], [-Ny * dh / 2, Ny * dh / 2])
    d = fdfd.Device(g, W200)
    fdfd.setup_eps_r(d, lambda x, y: abs(y) <= 0.15, 12.0)
    fdfd.setup_src(d, fdfd.Point(0.6, 0.0), fdfd.XHAT)
    return d


def oracle_fields(d):
    g = d.grid
    go = O.Grid2D(0.02, list(g.Npml), [g.bounds[0][0], g.bounds[1][0]], [g.bounds[0][1], g.bounds[1][1]])
    do = O.Device(go, [d.omega[0]])
    do.eps_r[:] = d.eps_r
    do.src[:] = d.src
    return O.solve(do, O.TM)["data"]


@pytest.mark.parametrize("nslabs", [1, 2, 4])
def test_slab_vs_oracle_waveguide(fdfd, nslabs):
    from fdfd_jl_b200 import slab
    d = waveguide(fdfd, 192, 128)
    f, infos = slab.solve_slabs_threads(d, nslabs)
    for i in infos:
        assert i["flag"] == 0 and i["relres"] <= RES_TOL
        assert i["iters"] == infos[0]["iters"]          # every slab took the same decisions
    assert rel(f.data, oracle_fields(d)) <= FIELD_TOL


def test_slab_vs_single_gpu_synthetic(fdfd):
    """512^2 synthetic map (bench workload at reduced size), 2 and 4 slabs against the single-GPU solve."""
    from fdfd_jl_b200 import slab, workloads
    d = workloads.synthetic_tm_device(fdfd, 512, 512, density=1.0 / 160.0)
    ref = fdfd.solve(d, fdfd.TM)
    assert ref.info["flag"] == 0
    for k in (2, 4):
        f, infos = slab.solve_slabs_threads(d, k)
        assert infos[0]["flag"] == 0 and infos[0]["relres"] <= RES_TOL
        assert rel(f.data, ref.data) <= FIELD_TOL
        # same cycle as on one GPU up to the cut PML lines: the iteration count stays in the same range
        assert infos[0]["iters"] <= 2 * ref.info["iters"] + 20


def test_slab_source_in_one_slab_only(fdfd):
    """point source: the right-hand side is zero on all slabs but one"""
    from fdfd_jl_b200 import slab
    g = fdfd.Grid(0.02, [12, 12], [0.0, 2.56], [0.0, 2.56])
    d = fdfd.Device(g, W200)
    fdfd.setup_src(d, fdfd.Point(1.0, 0.5))
    f, infos = slab.solve_slabs_threads(d, 4)
    assert infos[0]["flag"] == 0 and infos[0]["relres"] <= RES_TOL
    assert rel(f.data, oracle_fields(d)) <= FIELD_TOL


def test_slab_bad_arguments(fdfd):
    from fdfd_jl_b200 import slab
    g = fdfd.Grid(0.02, [10, 10], [0.0, 2.0], [0.0, 2.02])   # Ny = 101
    d = fdfd.Device(g, W200)
    with pytest.raises(fdfd.FdfdError):
        slab.solve_slabs_threads(d, 2)
