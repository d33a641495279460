"""Round 2 GPU experiment 1: multilevel Krylov variants at 4096^2 (single stream) and the 4-frequency sweep with 4 concurrent
solves, against BiCGSTAB + multigrid (both with the corner fix of the PML line relaxation)."""
import os, sys, time, math, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import _lib, workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
what = sys.argv[2] if len(sys.argv) > 2 else "all"
d = wl.synthetic_tm_device(fdfd, n, n, density=1 / 160.)

def pack(spec, restart):
    k = [int(x) for x in spec.split(",")] + [0, 0, 0]
    return k[0] | (k[1] << 8) | (k[2] << 16) | (restart << 24)

def single(spec, restart, env=None):
    for k, v in (env or {}).items(): os.environ[k] = v
    t0 = time.time()
    try:
        p = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, solver=_lib.SOLVER_MLKRYLOV, ml_spec=pack(spec, restart), maxit=3000)
        p.set_source(d.src)
        i = p.solve(); cyc = p.ml_cycles(); h = p.history()
        print(f"n={n} ML {spec:>8s} restart {restart} {env or ''}: flag={i['flag']} outer={i['iters']} relres={i['relres']:.2e} solve={i['solve_ms']:.0f} ms "
              f"restarts={i['restarts']} cycles={cyc[:3]} wall={time.time()-t0:.1f}s hist@10/20/40/80={[f'{h[min(k, len(h)-1)]:.1e}' for k in (10, 20, 40, 80)]}", flush=True)
        p.close()
    except Exception as e:
        print(f"n={n} ML {spec} restart {restart}: FAILED {str(e)[-200:]}", flush=True)
    for k in (env or {}): del os.environ[k]

def sweep(tag, **kw):
    d4 = wl.synthetic_tm_device(fdfd, n, n, density=1 / 160.)
    d4.omega = [2 * math.pi * (200e12 + 0.5e12 * k) for k in range(4)]
    t0 = time.time()
    try:
        fs = fdfd.solve(d4, fdfd.TM, maxit=6000, concurrency=4, **kw)
        print(f"n={n} SWEEP4 {tag}: wall={time.time()-t0:.1f}s iters={[f.info['iters'] for f in fs]} krylov_ms={[round(f.info['solve_ms']) for f in fs]} "
              f"flags={[f.info['flag'] for f in fs]}", flush=True)
    except Exception as e:
        print(f"n={n} SWEEP4 {tag}: FAILED {str(e)[-300:]}", flush=True)

if what in ("all", "single"):
    single("6,8", 96)
    single("6,8", 96, {"FDFD_ML_L0CGS": "1"})
    single("6,8", 160, {"FDFD_ML_L0CGS": "1"})
    single("4,6", 96, {"FDFD_ML_L0CGS": "1"})
    single("8,8", 96, {"FDFD_ML_L0CGS": "1"})
    single("6,6", 96, {"FDFD_ML_L0CGS": "1"})
    single("4,4,6", 96, {"FDFD_ML_L0CGS": "1"})
    single("10", 96, {"FDFD_ML_L0CGS": "1"})
    single("6,8", 48, {"FDFD_ML_L0CGS": "1"})
if what in ("all", "sweep"):
    sweep("bicgstab")
    os.environ["FDFD_ML_L0CGS"] = "1"
    sweep("ml 6,8 r48", solver=_lib.SOLVER_MLKRYLOV, ml_spec=pack("6,8", 48))
    sweep("ml 6,8 r64", solver=_lib.SOLVER_MLKRYLOV, ml_spec=pack("6,8", 64))
    print(subprocess.run(["nvidia-smi", "--query-gpu=memory.used,memory.total", "--format=csv,noheader"], capture_output=True, text=True).stdout)
