"""Round 2 GPU experiment 4: does stream concurrency overlap the latency-bound multilevel solves?  4-frequency sweep, concurrency 1/2/4."""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import _lib, workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
os.environ["FDFD_ML_L0CGS"] = "1"
def pack(spec, restart):
    k = [int(x) for x in spec.split(",")] + [0, 0, 0]
    return k[0] | (k[1] << 8) | (k[2] << 16) | (restart << 24)
d4 = wl.synthetic_tm_device(fdfd, n, n, density=1 / 160.)
d4.omega = [2 * math.pi * (200e12 + 0.5e12 * k) for k in range(4)]
def sweep(tag, conc, **kw):
    t0 = time.time()
    try:
        fs = fdfd.solve(d4, fdfd.TM, maxit=6000, concurrency=conc, **kw)
        print(f"n={n} SWEEP4 {tag} conc={conc}: wall={time.time()-t0:.1f}s iters={[f.info['iters'] for f in fs]} krylov_ms={[round(f.info['solve_ms']) for f in fs]} "
              f"setup_ms={[round(f.info['setup_ms']) for f in fs]} total_ms={[round(f.info['total_ms']) for f in fs]}", flush=True)
    except Exception as e:
        print(f"n={n} SWEEP4 {tag}: FAILED {str(e)[-300:]}", flush=True)
for pre in ("0", "1"):
    os.environ["FDFD_ML_PREALLOC"] = pre
    for conc in (1, 2, 4):
        sweep(f"mlF 6,6 r40 prealloc={pre}", conc, solver=_lib.SOLVER_MLKRYLOV, ml_spec=pack("6,6", 40), mg_cycle=1)
for conc in (1, 2, 4):
    sweep("bicg wd3", conc, mg_wdepth=3)
