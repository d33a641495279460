"""world_size-2 gloo test (CPU) of the N>1 path: omega-sweep sharding, gather order, no item lost or duplicated."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import fdfd_jl_b200 as fdfd
    from importlib import import_module
    sweep = import_module("fdfd_jl_b200.sweep")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = fdfd.Grid(0.1, [2, 2], [0, 1.0], [0, 1.0])
    ws = [1e15 + 1e13 * k for k in range(7)]
    d = fdfd.Device(g, ws)

    def fake_solve(d1, pol):  # stands in for the GPU solve: result identifies (omega, rank)
        assert len(d1.omega) == 1
        return {"omega": d1.omega[0], "rank": rank, "data": np.full((2, 2), d1.omega[0])}

    out = sweep.solve_sweep(d, fdfd.TM, fake_solve, rank=rank, world=world)
    if rank == 0:
        q.put([(o["omega"], o["rank"]) for o in out])
    dist.barrier()
    dist.destroy_process_group()


def test_sweep_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ws = [1e15 + 1e13 * k for k in range(7)]
    assert [w for w, _ in res] == ws                      # gathered in sweep order, nothing lost or duplicated
    assert [r for _, r in res] == [k % 2 for k in range(7)]  # round-robin ownership


def test_shard_indices_partition():
    sys.path.insert(0, ROOT)
    from importlib import import_module
    import fdfd_jl_b200  # noqa: F401
    sweep = import_module("fdfd_jl_b200.sweep")
    for n in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            parts = [sweep.shard_indices(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sweep.shard_indices(4, 2, 2)
