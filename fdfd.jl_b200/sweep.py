"""omega / source sweeps sharded over the GPUs of one box: independent units, no data-path collective.
One process per GPU (torchrun).  Two schedules:
  * static  -- rank r owns items r, r+W, r+2W, ... (no coordination at all);
  * dynamic -- ranks pull chunks of `chunk` consecutive items from a shared counter (an atomic add on the process group's
    store, a few bytes per chunk; still no data-path collective).  The iteration count of a solve grows with frequency and is
    erratic on resonant maps (round 1, 8 GPUs: the rank that held the highest frequencies took 1.4x as long as rank 0), so a
    fixed sweep (BASELINE config 3: 64 frequencies over 8 GPUs) finishes when the unluckiest static share does; the queue
    evens that out.
Results are optionally gathered on rank 0.  The reference loop is `for i in eachindex(d.ω)` (src/solver/driven.jl:11),
independent per ω."""
from __future__ import annotations

import copy


def shard_indices(n_items: int, rank: int, world: int):
    """round-robin ownership of the sweep items"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_items, world))


class WorkQueue:
    """shared counter over the sweep items: `next_chunk()` returns the next unclaimed index range (empty when exhausted).
    `store` is any torch.distributed store (TCPStore / the default group's store); its `add` is atomic across ranks."""

    def __init__(self, store, n_items: int, chunk: int = 1, key: str = "fdfd_sweep_queue"):
        if chunk < 1:
            raise ValueError("chunk must be >= 1")
        self.store, self.n, self.chunk, self.key = store, n_items, chunk, key

    def next_chunk(self):
        hi = self.store.add(self.key, self.chunk)   # value AFTER the add
        lo = hi - self.chunk
        return range(min(lo, self.n), min(hi, self.n))


_queue_serial = 0


def solve_sweep(device, pol, solve_fn, rank=0, world=1, gather=True, group=None, schedule="static", chunk=1, store=None):
    """Solve every frequency of `device.omega`, sharded over `world` ranks.  `solve_fn(device_with_some_omegas, pol)` is the
    solve of one chunk (fdfd.solve on a GPU rank; it returns a field, or a list of fields for a chunk of several
    frequencies -- a chunk of 4 lets the library overlap them on its worker streams).  schedule: "static" | "dynamic".
    Returns the full list of fields on rank 0 (others: their own (index, field) pairs) when gather=True, else the local pairs."""
    global _queue_serial
    n = len(device.omega)
    local = []

    def run(idx):
        if not len(idx):
            return
        d1 = copy.copy(device)
        d1.omega = [device.omega[i] for i in idx]
        out = solve_fn(d1, pol)
        outs = out if isinstance(out, list) and len(idx) > 1 else [out]
        if len(outs) != len(idx):
            raise RuntimeError("solve_fn returned a different number of fields than frequencies")
        local.extend(zip(idx, outs))

    if schedule == "dynamic" and world > 1:
        import torch.distributed as dist
        if store is None:
            store = dist.distributed_c10d._get_default_store()
        _queue_serial += 1   # every rank calls solve_sweep the same number of times: same key on every rank
        q = WorkQueue(store, n, chunk, key=f"fdfd_sweep_queue_{_queue_serial}")
        while True:
            r = q.next_chunk()
            if not len(r):
                break
            run(list(r))
    elif schedule in ("static", "dynamic"):
        mine = shard_indices(n, rank, world)
        for k in range(0, len(mine), chunk):
            run(mine[k:k + chunk])
    else:
        raise ValueError(f"unknown schedule {schedule!r}")
    if world == 1 or not gather:
        return [f for _, f in local] if world == 1 else local
    import torch.distributed as dist
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
    if rank != 0:
        return local
    out = [None] * len(device.omega)
    for part in bucket:
        for i, f in part:
            out[i] = f
    return out
