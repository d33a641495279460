"""One large grid split into row slabs over several GPUs (BASELINE config 5): the sharded form of
solve(d::Device, TM) (src/solver/driven.jl:4-59).  The reference has no parallel path; see csrc/slab.cu for the scheme.

Two ways to run it:
  * one process per GPU (torchrun): `comm = SlabComm.nccl(ctx, rank, world)` then `solve_slab(d, comm, ctx)` on every
    rank.  The halo exchange and the Krylov allreduce run inside the library over NCCL; torch.distributed is only used
    to broadcast the 128-byte NCCL id.
  * one process, several slabs (tests; also several GPUs of one box without torchrun): `solve_slabs_threads(d, nslabs)`.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import Context, Info, check, default_opts, lib, ptr


def slab_rows(grid, nranks: int, rank: int):
    """(y0, nrows) owned by `rank`: global rows [y0, y0+nrows).  Host-only."""
    y0, n = C.c_int64(), C.c_int64()
    gc = grid.as_c()
    check(lib().fdfd_slab_rows(C.byref(gc), int(nranks), int(rank), C.byref(y0), C.byref(n)), None)
    return int(y0.value), int(n.value)


class SlabComm:
    """Handle of a slab communicator (fdfd_comm*)."""

    def __init__(self, handle, nranks, rank, kind, keep=None):
        self._h, self.nranks, self.rank, self.kind, self._keep = handle, nranks, rank, kind, keep

    @classmethod
    def nccl(cls, ctx: Context, rank: int, world: int, group=None):
        """One process per GPU.  Rank 0 creates the NCCL id; torch.distributed (any backend) broadcasts it."""
        import torch
        import torch.distributed as dist
        buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
        if rank == 0:
            check(lib().fdfd_comm_unique_id(buf), None)
        if world > 1:
            box = [bytes(buf.raw)]
            dist.broadcast_object_list(box, src=0, group=group)
            buf = C.create_string_buffer(box[0], _lib.COMM_ID_BYTES)
        h = C.c_void_p()
        check(lib().fdfd_comm_create_nccl(ctx.handle, int(world), int(rank), buf, C.byref(h)), ctx.handle)
        return cls(h, world, rank, _lib.COMM_NCCL)

    @classmethod
    def threads(cls, nranks: int):
        """All slabs in this process: returns one communicator per rank (use one host thread per rank)."""
        g = C.c_void_p()
        check(lib().fdfd_comm_group_create(int(nranks), C.byref(g)), None)
        grp = _Group(g)
        out = []
        for r in range(nranks):
            h = C.c_void_p()
            check(lib().fdfd_comm_create_threads(g, r, C.byref(h)), None)
            out.append(cls(h, nranks, r, _lib.COMM_THREADS, keep=grp))
        return out

    @property
    def handle(self):
        return self._h

    def stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        lib().fdfd_comm_stats(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"exchanges": a.value, "allreduces": b.value, "bytes_sent": c.value}

    def close(self):
        if self._h:
            lib().fdfd_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Group:
    def __init__(self, h):
        self._h = h

    def __del__(self):
        try:
            if self._h:
                lib().fdfd_comm_group_destroy(self._h)
                self._h = None
        except Exception:
            pass


def solve_slab_rows(grid, omega, eps_rows, src_rows, comm: SlabComm, ctx: Context, **kw):
    """Lowest level: this rank's rows in, this rank's (Nx, nrows, 3) field rows out.  eps_rows / src_rows are (Nx, nrows)
    arrays (or device pointers).  Collective over `comm`."""
    o = kw.pop("opts", None) or default_opts(**kw)
    Nx, _ = grid.N
    _, n = slab_rows(grid, comm.nranks, comm.rank)
    if not isinstance(eps_rows, (int, np.integer)):
        eps_rows = _lib.as_c128(eps_rows, (Nx, n))
    if not isinstance(src_rows, (int, np.integer)):
        src_rows = _lib.as_c128(src_rows, (Nx, n))
    fields = np.empty((Nx, n, 3), dtype=np.complex128, order="F")
    info = Info()
    gc = grid.as_c()
    code = lib().fdfd_solve_driven_slab(ctx.handle, comm.handle, C.byref(gc), float(omega), ptr(eps_rows), ptr(src_rows),
                                        C.byref(o), ptr(fields), C.byref(info))
    check(code, ctx.handle)
    return fields, info.asdict()


def solve_slab(d, comm: SlabComm, ctx: Context, **kw):
    """solve(d, TM) with the grid of `d` split into comm.nranks row slabs; returns this rank's rows of
    FieldTM.data, shape (Nx, nrows, 3), and the solve info.  `d` is the full device on every rank (only the owned rows of
    eps_r / src are read)."""
    if len(d.omega) != 1:
        raise ValueError("solve_slab takes a single frequency")
    y0, n = slab_rows(d.grid, comm.nranks, comm.rank)
    return solve_slab_rows(d.grid, d.omega[0], d.eps_r[:, y0:y0 + n], d.src[:, y0:y0 + n], comm, ctx, **kw)


def solve_slabs_threads(d, nslabs: int, devices=None, **kw):
    """All slabs from one process, one host thread (own context and stream) per slab; `devices[r]` is the CUDA ordinal of
    slab r (default: all on device 0).  Returns (FieldTM, [info per slab])."""
    from . import FieldTM
    comms = SlabComm.threads(nslabs)
    devices = devices or [0] * nslabs
    ctxs = [Context(dev) for dev in devices]
    Nx, Ny = d.grid.N
    data = np.empty((Nx, Ny, 3), dtype=np.complex128, order="F")
    infos, errs = [None] * nslabs, [None] * nslabs

    def work(r):
        try:
            y0, n = slab_rows(d.grid, nslabs, r)
            f, info = solve_slab(d, comms[r], ctxs[r], **dict(kw))
            data[:, y0:y0 + n, :] = f
            infos[r] = info
        except Exception as e:  # noqa: BLE001 - reported below
            errs[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(nslabs)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in comms:
        c.close()
    for c in ctxs:
        c.close()
    for e in errs:
        if e is not None:
            raise e
    return FieldTM(d.grid, d.omega[0], data, infos[0]), infos


def _run_threads(nslabs, devices, work):
    """one host thread (own context and stream) per slab; re-raises the first error"""
    comms = SlabComm.threads(nslabs)
    devices = devices or [0] * nslabs
    ctxs = [Context(dev) for dev in devices]
    errs = [None] * nslabs

    def run(r):
        try:
            work(r, comms[r], ctxs[r])
        except Exception as e:  # noqa: BLE001 - reported below
            errs[r] = e

    th = [threading.Thread(target=run, args=(r,)) for r in range(nslabs)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in comms:
        c.close()
    for c in ctxs:
        c.close()
    for e in errs:
        if e is not None:
            raise e


# ---- slab-sharded modulated / eigenfrequency solves (csrc/slab_multi.cu) ----
def solve_modulated_slab_rows(grid, omega, Omega, nsidebands, sharedpml, eps_rows, deps_rows, src_rows, comm: SlabComm,
                              ctx: Context, **kw):
    """This rank's rows of solve(d::ModulatedDevice) (modulation.jl:35-119): returns fields (Nx, nrows, 3, nf), sideband
    -ns first, and the solve info.  Collective over `comm`."""
    o = kw.pop("opts", None) or default_opts(**kw)
    Nx, _ = grid.N
    _, n = slab_rows(grid, comm.nranks, comm.rank)
    nf = 2 * int(nsidebands) + 1
    eps_rows = _lib.as_c128(eps_rows, (Nx, n)); deps_rows = _lib.as_c128(deps_rows, (Nx, n)); src_rows = _lib.as_c128(src_rows, (Nx, n))
    fields = np.empty((Nx, n, 3, nf), dtype=np.complex128, order="F")
    info = Info()
    gc = grid.as_c()
    code = lib().fdfd_solve_modulated_slab(ctx.handle, comm.handle, C.byref(gc), float(omega), float(Omega), int(nsidebands),
                                           int(bool(sharedpml)), ptr(eps_rows), ptr(deps_rows), ptr(src_rows), C.byref(o),
                                           ptr(fields), C.byref(info))
    check(code, ctx.handle)
    return fields, info.asdict()


def solve_modulated_slabs_threads(d, nslabs: int, devices=None, **kw):
    """solve(d::ModulatedDevice) with the grid cut into `nslabs` row slabs, all from one process (tests).  Returns
    ([FieldTM per sideband], [info per slab]); single frequency."""
    from . import FieldTM
    if len(d.omega) != 1:
        raise ValueError("the slab modulated solve takes a single frequency")
    Nx, Ny = d.grid.N
    nf = 2 * d.nsidebands + 1
    data = np.empty((Nx, Ny, 3, nf), dtype=np.complex128, order="F")
    infos = [None] * nslabs

    def work(r, comm, ctx):
        y0, n = slab_rows(d.grid, nslabs, r)
        f, info = solve_modulated_slab_rows(d.grid, d.omega[0], d.Omega, d.nsidebands, d.sharedpml, d.eps_r[:, y0:y0 + n],
                                            d.deps_r[:, y0:y0 + n], d.src[:, y0:y0 + n], comm, ctx, **dict(kw))
        data[:, y0:y0 + n, :, :] = f
        infos[r] = info

    _run_threads(nslabs, devices, work)
    omegan = d.omega[0] + d.Omega * np.arange(-d.nsidebands, d.nsidebands + 1)
    return [FieldTM(d.grid, omegan[j], data[:, :, :, j], infos[0]) for j in range(nf)], infos


def eigenfrequency_slab_rows(grid, omega0, nev, eps_rows, comm: SlabComm, ctx: Context, which="LM", ncv=0, want_fields=True, **kw):
    """This rank's rows of eigenfrequency(d, TM, nev; which) (eigen.jl:69-96): (ω[nev], fields (Nx, nrows, 3, nev) or None,
    info).  Collective over `comm`; ω is identical on every rank."""
    o = kw.pop("opts", None) or default_opts(**kw)
    Nx, _ = grid.N
    _, n = slab_rows(grid, comm.nranks, comm.rank)
    eps_rows = _lib.as_c128(eps_rows, (Nx, n))
    om = np.empty(nev, dtype=np.complex128)
    fields = np.empty((Nx, n, 3, nev), dtype=np.complex128, order="F") if want_fields else None
    info = Info()
    gc = grid.as_c()
    code = lib().fdfd_eigenfrequency_slab(ctx.handle, comm.handle, C.byref(gc), _lib.TM, float(omega0), int(nev),
                                          _lib.WHICH[which], int(ncv), ptr(eps_rows), C.byref(o), ptr(om), ptr(fields),
                                          C.byref(info))
    check(code, ctx.handle)
    return om, fields, info.asdict()


def eigenfrequency_slabs_threads(d, nev, nslabs: int, which="LM", ncv=0, devices=None, **kw):
    """eigenfrequency(d, TM, nev) with the grid cut into `nslabs` row slabs, all from one process (tests).
    Returns (ω, [FieldTM], [info per slab])."""
    from . import FieldTM
    Nx, Ny = d.grid.N
    data = np.empty((Nx, Ny, 3, nev), dtype=np.complex128, order="F")
    oms, infos = [None] * nslabs, [None] * nslabs

    def work(r, comm, ctx):
        y0, n = slab_rows(d.grid, nslabs, r)
        om, f, info = eigenfrequency_slab_rows(d.grid, d.omega[0], nev, d.eps_r[:, y0:y0 + n], comm, ctx, which=which, ncv=ncv, **dict(kw))
        data[:, y0:y0 + n, :, :] = f
        oms[r], infos[r] = om, info

    _run_threads(nslabs, devices, work)
    for om in oms[1:]:
        if not np.array_equal(om, oms[0]):
            raise RuntimeError("slab eigenfrequencies differ between ranks")
    return oms[0], [FieldTM(d.grid, oms[0][i], data[:, :, :, i], infos[0]) for i in range(nev)], infos


def gather_field(grid, omega, rows_data, rank, world, info=None, group=None):
    """Assemble the full FieldTM on rank 0 from every rank's (Nx, nrows, 3) rows (torch.distributed gather; any backend).
    Other ranks get None.  Only for grids whose (Nx,Ny,3) field fits one host; large runs keep the rows sharded."""
    from . import FieldTM
    if world == 1:
        return FieldTM(grid, omega, np.asfortranarray(rows_data), info)
    import torch.distributed as dist
    bucket = [None] * world if rank == 0 else None
    dist.gather_object((rank, np.ascontiguousarray(rows_data)), bucket, dst=0, group=group)
    if rank != 0:
        return None
    Nx, Ny = grid.N
    data = np.empty((Nx, Ny, 3), dtype=np.complex128, order="F")
    for r, part in bucket:
        y0, n = slab_rows(grid, world, r)
        data[:, y0:y0 + n, :] = part
    return FieldTM(grid, omega, data, info)
