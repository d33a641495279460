"""omega / source sweeps sharded over the GPUs of one box: independent units, no data-path collective.
One process per GPU (torchrun); rank r owns items r, r+W, r+2W, ...; results are optionally gathered on rank 0.
The reference loop is `for i in eachindex(d.ω)` (src/solver/driven.jl:11), independent per ω."""
from __future__ import annotations

import copy


def shard_indices(n_items: int, rank: int, world: int):
    """round-robin ownership of the sweep items"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_items, world))


def solve_sweep(device, pol, solve_fn, rank=0, world=1, gather=True, group=None):
    """Solve every frequency of `device.omega`, sharded over `world` ranks.  `solve_fn(device_with_one_omega, pol)` is the
    single-frequency solve (fdfd.solve on a GPU rank).  Returns the full list of fields on rank 0 (others: their own
    (index, field) pairs) when gather=True, else the local pairs."""
    mine = shard_indices(len(device.omega), rank, world)
    local = []
    for i in mine:
        d1 = copy.copy(device)
        d1.omega = [device.omega[i]]
        local.append((i, solve_fn(d1, pol)))
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
