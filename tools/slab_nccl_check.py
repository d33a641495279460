"""torchrun check of the NCCL transport of the slab solve: every rank solves its slab over NCCL, then solves the whole
grid alone on its own GPU and compares its rows.   torchrun --nproc-per-node N tools/slab_nccl_check.py [grid]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import slab, workloads

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = fdfd.Context(local)
comm = slab.SlabComm.nccl(ctx, rank, world)
d = workloads.synthetic_tm_device(fdfd, n, n, density=1.0 / 160.0)
y0, nr = slab.slab_rows(d.grid, world, rank)
for use_graph in (0, 1):
    t0 = time.time()
    f, info = slab.solve_slab(d, comm, ctx, use_graph=use_graph)
    t1 = time.time()
    if use_graph == 0:
        ref = fdfd.solve(d, fdfd.TM, ctx=ctx)
    err = float(np.linalg.norm(f - ref.data[:, y0:y0 + nr, :]) / np.linalg.norm(ref.data[:, y0:y0 + nr, :]))
    print(f"[rank {rank}/{world}] n={n} graph={use_graph} iters={info['iters']} (single {ref.info['iters']}) relres={info['relres']:.2e} flag={info['flag']} "
          f"solve_ms={info['solve_ms']:.0f} (single {ref.info['solve_ms']:.0f}) rel_vs_single={err:.2e} wall={t1-t0:.1f}s stats={comm.stats()}", flush=True)
    assert info["flag"] == 0 and info["relres"] <= 1e-10 and err <= 1e-6
comm.close()
if world > 1:
    dist.destroy_process_group()
print(f"[rank {rank}] OK")
