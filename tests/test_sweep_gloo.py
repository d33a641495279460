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


def _worker_dynamic(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import time
    import torch.distributed as dist
    import fdfd_jl_b200 as fdfd
    from importlib import import_module
    sweep = import_module("fdfd_jl_b200.sweep")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = fdfd.Grid(0.1, [2, 2], [0, 1.0], [0, 1.0])
    ws = [1e15 + 1e13 * k for k in range(13)]
    d = fdfd.Device(g, ws)

    def fake_solve(d1, pol):   # rank 0 is the slow GPU: the queue must hand most chunks to rank 1
        time.sleep(0.25 if rank == 0 else 0.01)
        return [{"omega": w, "rank": rank} for w in d1.omega] if len(d1.omega) > 1 else {"omega": d1.omega[0], "rank": rank}

    for _ in range(2):   # two sweeps in a row: the queue key must not be reused
        out = sweep.solve_sweep(d, fdfd.TM, fake_solve, rank=rank, world=world, schedule="dynamic", chunk=2)
        dist.barrier()
    if rank == 0:
        q.put([(o["omega"], o["rank"]) for o in out])
    dist.barrier()
    dist.destroy_process_group()


def test_sweep_dynamic_queue_world2():
    """dynamic schedule (shared counter on the group's store): nothing lost or duplicated, sweep order kept, and the fast rank
    takes more chunks than the slow one"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_dynamic, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ws = [1e15 + 1e13 * k for k in range(13)]
    assert [w for w, _ in res] == ws
    owners = [r for _, r in res]
    assert owners.count(1) > owners.count(0) >= 1
    for k in range(0, 12, 2):           # chunks of 2 consecutive frequencies stay on one rank
        assert owners[k] == owners[k + 1]


def test_work_queue_chunks():
    sys.path.insert(0, ROOT)
    from importlib import import_module
    import fdfd_jl_b200  # noqa: F401
    sweep = import_module("fdfd_jl_b200.sweep")

    class Store:   # the one method the queue uses
        def __init__(self): self.v = {}
        def add(self, k, n): self.v[k] = self.v.get(k, 0) + n; return self.v[k]

    st = Store()
    qa, qb = sweep.WorkQueue(st, 10, 4), sweep.WorkQueue(st, 10, 4)   # two ranks, one counter
    got = [list(qa.next_chunk()), list(qb.next_chunk()), list(qa.next_chunk()), list(qb.next_chunk()), list(qa.next_chunk())]
    assert got == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9], [], []]
    with pytest.raises(ValueError):
        sweep.WorkQueue(st, 10, 0)


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
