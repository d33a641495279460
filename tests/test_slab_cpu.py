"""CPU tests of the host side of the row-slab sharded solve: slab ownership, the row-range workload
generator, the world-size-2 gloo path of the NCCL-id broadcast and of the field gather, and that nothing on this path
computes without a GPU (no CPU fallback)."""
import os
import socket

import numpy as np
import pytest


def test_slab_rows_partition(fdfd):
    from fdfd_jl_b200 import slab
    g = fdfd.Grid(0.02, [15, 15], [0.0, 20.48], [0.0, 40.96])   # 1024 x 2048
    for world in (1, 2, 4, 8):
        got = [slab.slab_rows(g, world, r) for r in range(world)]
        assert got[0][0] == 0 and all(n == g.N[1] // world for _, n in got)
        assert all(got[r][0] == got[r - 1][0] + got[r - 1][1] for r in range(1, world))
        assert got[-1][0] + got[-1][1] == g.N[1]
    with pytest.raises(fdfd.FdfdError):
        slab.slab_rows(g, 3, 0)        # 2048 rows do not split into 3 equal slabs
    with pytest.raises(fdfd.FdfdError):
        slab.slab_rows(g, 2, 2)


def test_workload_rows_match_full_map(fdfd):
    from fdfd_jl_b200 import slab, workloads
    d = workloads.synthetic_tm_device(fdfd, 256, 384, density=1.0 / 40.0)
    for world in (2, 4):
        for r in range(world):
            y0, n = slab.slab_rows(d.grid, world, r)
            g, w, eps, src = workloads.synthetic_tm_device(fdfd, 256, 384, density=1.0 / 40.0, rows=(y0, n))
            assert g.N == d.grid.N and w == d.omega[0]
            assert np.array_equal(eps, d.eps_r[:, y0:y0 + n]) and np.array_equal(src, d.src[:, y0:y0 + n])


def test_no_cpu_fallback_for_slab(fdfd):
    """without a CUDA device the thread communicator cannot be created and the NCCL one has no context to bind to"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fdfd_jl_b200 import slab
    with pytest.raises(fdfd.FdfdError):
        slab.SlabComm.threads(2)
    with pytest.raises(fdfd.FdfdError):
        fdfd.Context(0)


def test_no_cpu_fallback_for_slab_modulated_eigen_and_multilevel(fdfd):
    """the slab-sharded modulated / eigenfrequency entry points and the multilevel Krylov solver fail loudly without a GPU"""
    import math
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fdfd_jl_b200 import slab
    g = fdfd.Grid(0.02, [10, 10], [0.0, 2.56], [0.0, 2.56])
    md = fdfd.ModulatedDevice(g, 2 * math.pi * 200e12, 4.5e14, 1)
    with pytest.raises(fdfd.FdfdError):
        slab.solve_modulated_slabs_threads(md, 2)
    with pytest.raises(fdfd.FdfdError):
        slab.eigenfrequency_slabs_threads(fdfd.Device(g, 2 * math.pi * 200e12), 2, 2)
    with pytest.raises(fdfd.FdfdError):
        fdfd.solve(fdfd.Device(g, 2 * math.pi * 200e12), fdfd.TM, solver=fdfd._lib.SOLVER_MLKRYLOV)
    # NULL / out-of-range arguments of the new entry points are rejected before any device work
    import ctypes as C
    L = fdfd.lib()
    gc = g.as_c()
    assert L.fdfd_solve_modulated_slab(None, None, C.byref(gc), 1.0, 1.0, 1, 1, None, None, None, None, None, None) == fdfd._lib.ERR_ARG
    assert L.fdfd_eigenfrequency_slab(None, None, C.byref(gc), fdfd.TM, 1.0, 1, 0, 0, None, None, None, None, None) == fdfd._lib.ERR_ARG
    assert L.fdfd_problem_ml_cycles(None, None) == fdfd._lib.ERR_ARG
    assert L.fdfd_debug_ml_lsq(0, None, 1.0, None, None) == fdfd._lib.ERR_ARG
    assert L.fdfd_debug_ml_transfer(1, 1, 0, 1.0, None, None) == fdfd._lib.ERR_ARG


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import ctypes as C
    import torch.distributed as dist
    import fdfd_jl_b200 as fdfd
    from fdfd_jl_b200 import slab, _lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) the NCCL id made on rank 0 reaches every rank unchanged (what SlabComm.nccl does before ncclCommInitRank)
        buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
        if rank == 0:
            _lib.check(_lib.lib().fdfd_comm_unique_id(buf), None)
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0)
        ids = [None] * world
        dist.all_gather_object(ids, box[0])
        # (2) every rank's rows of a known field land in the right place on rank 0
        g = fdfd.Grid(0.05, [4, 4], [0.0, 1.6], [0.0, 3.2])    # 32 x 64
        Nx, Ny = g.N
        full = (np.arange(Nx * Ny * 3, dtype=np.float64).reshape((Nx, Ny, 3), order="F") * (1 + 0.5j))
        y0, n = slab.slab_rows(g, world, rank)
        f = slab.gather_field(g, 1.0, full[:, y0:y0 + n, :], rank, world)
        ok_field = (f is None) if rank != 0 else bool(np.array_equal(f.data, full))
        q.put((rank, len(set(ids)) == 1 and len(ids[0]) == _lib.COMM_ID_BYTES and any(ids[0]), ok_field))
    finally:
        dist.destroy_process_group()


def test_slab_host_path_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True), (1, True, True)]
