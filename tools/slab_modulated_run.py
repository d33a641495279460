"""BASELINE config 5, second half: solve(d::ModulatedDevice) (modulation.jl:35-119) of a long modulated waveguide split into row slabs
over the ranks (fdfd_solve_modulated_slab, NCCL).  The device is the notebook's Example 3 (cells 17-25: eps = 12.25 guide of width
a = 0.2202 um, travelling-wave modulation exp(i q x) with q = 2.9263 / um in the lower half of the guide, Omega = 4.541e14, ns = 1)
stretched to Nx x Ny cells of dh = 0.01 um; x-normal line source across the guide.  Prints one JSON line on rank 0.
    torchrun --nproc-per-node 8 tools/slab_modulated_run.py 8192 2048"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import slab

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Nx, Ny = int(sys.argv[1]), int(sys.argv[2])
dh, a, q = 0.01, 0.2202, 2.9263
w, Om = 2 * math.pi * 1.939e14, 4.541e14
g = fdfd.Grid(dh, [15, 15], [0.0, Nx * dh], [-Ny * dh / 2, Ny * dh / 2])
assert g.N == (Nx, Ny)
ctx = fdfd.Context(local)
comm = slab.SlabComm.nccl(ctx, rank, world)
y0, nr = slab.slab_rows(g, world, rank)
xs = fdfd.xc(g)[:, None]; ys = fdfd.yc(g)[None, y0:y0 + nr]
eps = np.ones((Nx, nr), dtype=np.complex128); eps[np.broadcast_to((ys >= -a / 2) & (ys <= a / 2), eps.shape)] = 12.25
Lx = Nx * dh
deps = np.zeros((Nx, nr), dtype=np.complex128)
mod = (xs >= 0.1 * Lx) & (xs <= 0.9 * Lx) & (ys >= -a / 2) & (ys <= 0)
deps[mod] = np.broadcast_to(np.exp(1j * q * xs), deps.shape)[mod]
src = np.zeros((Nx, nr), dtype=np.complex128)
src[25, :] = np.where(np.abs(ys[0]) <= 2 * a, 1j, 0)
t0 = time.time()
try:
    f, info = slab.solve_modulated_slab_rows(g, w, Om, 1, True, eps, deps, src, comm, ctx, maxit=20000)
    err = None
except fdfd.FdfdError as e:
    f, info, err = None, None, str(e)[-300:]
wall = time.time() - t0
if rank == 0:
    rec = {"what": "solve(d::ModulatedDevice) on row slabs (fdfd_solve_modulated_slab), NCCL", "grid": [Nx, Ny], "sidebands": 3, "n_gpus": world,
           "unknowns": 3 * Nx * Ny, "rows_per_slab": nr, "wall_s": wall, "error": err}
    if info:
        rec.update({"converged": info["flag"] == 0 and info["relres"] <= 1e-10, "iters": info["iters"], "relres": info["relres"], "krylov_ms": info["solve_ms"],
                    "setup_ms": info["setup_ms"], "ms_per_iteration": info["solve_ms"] / max(1, info["iters"]), "mg_levels": info["mg_levels"],
                    "sideband_energy_rows0": [float(np.linalg.norm(f[:, :, 0, j])) for j in range(3)], "comm": comm.stats()})
    print(json.dumps(rec), flush=True)
comm.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
