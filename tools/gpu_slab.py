"""GPU exploration of the slab-sharded solve on ONE GPU (thread transport): single-GPU solve vs 1/2/4 slabs."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import slab, workloads

def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))

sizes = [int(s) for s in sys.argv[1].split(",")] if len(sys.argv) > 1 else [512]
nsl = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4]
for n in sizes:
    d = workloads.synthetic_tm_device(fdfd, n, n, density=1.0 / 160.0)
    t0 = time.time(); f = fdfd.solve(d, fdfd.TM); t1 = time.time()
    print(f"n={n} single: iters={f.info['iters']} relres={f.info['relres']:.2e} solve_ms={f.info['solve_ms']:.0f} levels={f.info['mg_levels']} wall={t1-t0:.1f}s", flush=True)
    for k in nsl:
        try:
            t0 = time.time(); fs, infos = slab.solve_slabs_threads(d, k); t1 = time.time()
            i0 = infos[0]
            print(f"n={n} slabs={k}: iters={i0['iters']} relres={i0['relres']:.2e} flag={i0['flag']} solve_ms={i0['solve_ms']:.0f} levels={i0['mg_levels']} "
                  f"restarts={i0['restarts']} rel_vs_single={rel(fs.data, f.data):.2e} relEz={rel(fs.data[:,:,0], f.data[:,:,0]):.2e} wall={t1-t0:.1f}s", flush=True)
        except Exception as e:
            print(f"n={n} slabs={k}: FAILED {e}", flush=True)
