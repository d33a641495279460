"""Worker of tests/test_gpu_slab_nccl.py (one process per GPU under torch.distributed.run): the NCCL transport of the row-slab
sharded solves (csrc/comm.cu ring send/recv + allgather + allreduce) -- driven, modulated and eigenfrequency -- against the
single-GPU solves of the same devices and, for the driven solve, the oracle's direct solve.  Not collected by pytest (leading _)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import fdfd_jl_b200 as fdfd  # noqa: E402
from fdfd_jl_b200 import slab, workloads  # noqa: E402

FIELD_TOL, RES_TOL, EIG_TOL = 1e-6, 1e-10, 1e-8


def rel(a, b):
    return float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b)))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = fdfd.Context(local)
    comm = slab.SlabComm.nccl(ctx, rank, world)

    # ---- driven TM, the bench map at 512^2: slab rows == single-GPU rows == oracle rows
    d = workloads.synthetic_tm_device(fdfd, 512, 512, density=1.0 / 160.0)
    y0, nr = slab.slab_rows(d.grid, world, rank)
    ref = fdfd.solve(d, fdfd.TM, ctx=ctx)
    for use_graph in (1, 0):
        f, info = slab.solve_slab(d, comm, ctx, use_graph=use_graph)
        assert info["flag"] == 0 and info["relres"] <= RES_TOL, info
        assert rel(f, ref.data[:, y0:y0 + nr, :]) <= FIELD_TOL
    st = comm.stats()
    assert st["exchanges"] > 0 and st["allreduces"] > 0 and st["bytes_sent"] > 0, st          # the NCCL path really ran
    if rank == 0:
        from oracle import fdfd_oracle as O
        go = O.Grid2D(0.02, [15, 15], [0.0, 512 * 0.02], [0.0, 512 * 0.02])
        do = O.Device(go, [d.omega[0]]); do.eps_r[:] = d.eps_r; do.src[:] = d.src
        fo = O.solve(do, O.TM)["data"]
        assert rel(f, fo[:, y0:y0 + nr, :]) <= FIELD_TOL

    # ---- modulated (modulation.jl:35-119) on slabs over NCCL against the single-GPU solve
    w, Om, a, q = 2 * math.pi * 1.939e14, 4.541e14, 0.2202, 2.9263
    g = fdfd.Grid(0.02, [15, 15], [0.0, 5.0], [-1.28, 1.28])
    dm = fdfd.ModulatedDevice(g, w, Om, 1)
    fdfd.setup_eps_r(dm, lambda x, y: -a / 2 <= y <= a / 2, 12.25)
    fdfd.setup_deps_r(dm, lambda x, y: (1 <= x <= 4.0) and (-a / 2 <= y <= 0), lambda x, y: np.exp(1j * q * x))
    fdfd.add_mode(dm, fdfd.Mode(fdfd.TM, fdfd.XHAT, 3.5, fdfd.Point(0.2, 0), 4 * a))
    refm = fdfd.solve(dm, ctx=ctx)[0]      # also launches the mode source into dm.src
    y0, nr = slab.slab_rows(g, world, rank)
    fm, info = slab.solve_modulated_slab_rows(g, w, Om, 1, True, dm.eps_r[:, y0:y0 + nr], dm.deps_r[:, y0:y0 + nr], dm.src[:, y0:y0 + nr], comm, ctx)
    assert info["flag"] == 0 and info["relres"] <= RES_TOL, info
    scale = max(np.linalg.norm(r.data) for r in refm)
    for j in range(3):
        assert np.linalg.norm(fm[:, :, :, j] - refm[j].data[:, y0:y0 + nr, :]) <= FIELD_TOL * scale

    # ---- eigenfrequency (eigen.jl:69-96) on slabs over NCCL: sharded Arnoldi basis, allreduced dots, Krylov-Schur restarts
    ge = fdfd.Grid(0.025, [15, 15], [-2.0, 2.0], [-2.0, 2.0])      # 160 x 160
    de = fdfd.Device(ge, 2 * math.pi * 200e12)
    fdfd.setup_eps_r(de, [fdfd.Cylinder((0, 0), 0.8, 1.0), fdfd.Cylinder((0, 0), 1.0, 12.25)])
    om_ref, _ = fdfd.eigenfrequency(de, fdfd.TM, 6, which="LM", ctx=ctx)
    y0, nr = slab.slab_rows(ge, world, rank)
    out = slab.eigenfrequency_slab_rows(ge, de.omega[0], 4, de.eps_r[:, y0:y0 + nr], comm, ctx, which="LM", want_fields=False)
    om = out[0]
    left = list(om_ref)
    for z in om:
        k = int(np.argmin([abs(z - r) for r in left]))
        assert abs(z - left[k]) / abs(left[k]) <= EIG_TOL
        left.pop(k)

    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"[rank {rank}] NCCL_SLAB_OK", flush=True)


if __name__ == "__main__":
    main()
