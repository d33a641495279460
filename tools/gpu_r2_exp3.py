"""Round 2 GPU experiment 3: stronger cycles (deeper W, F) for BiCGSTAB and the multilevel Krylov solver at 4096^2."""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import _lib, workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = wl.synthetic_tm_device(fdfd, n, n, density=1 / 160.)
os.environ["FDFD_ML_L0CGS"] = "1"
def pack(spec, restart):
    k = [int(x) for x in spec.split(",")] + [0, 0, 0]
    return k[0] | (k[1] << 8) | (k[2] << 16) | (restart << 24)
def run(tag, env=None, **kw):
    for k, v in (env or {}).items(): os.environ[k] = v
    try:
        p = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, maxit=4000, **kw)
        p.set_source(d.src)
        i = p.solve()
        cyc = p.ml_cycles() if kw.get("solver") == _lib.SOLVER_MLKRYLOV else ""
        print(f"n={n} {tag} {env or ''} {kw}: flag={i['flag']} iters={i['iters']} solve={i['solve_ms']:.0f} ms restarts={i['restarts']} launches={i['launches']} "
              f"ms/it={i['solve_ms']/max(1,i['iters']):.2f} {cyc}", flush=True)
        p.close()
    except Exception as e:
        print(f"n={n} {tag} {env} {kw}: FAILED {str(e)[-200:]}", flush=True)
    for k in (env or {}): del os.environ[k]
def ML(spec, restart=96, **kw): return dict(solver=_lib.SOLVER_MLKRYLOV, ml_spec=pack(spec, restart), **kw)
for wd in (3, 4, 5, 9):
    run("bicg", {}, mg_wdepth=wd)
run("bicg", {}, mg_wdepth=3, mg_nu2=2)
run("bicg", {}, mg_wdepth=3, mg_coarse_sweeps=1)
for spec in ("6,6", "6,8", "4,6", "8,8", "4,4", "3,6", "8", "12"):
    run("mlF " + spec, {}, **ML(spec, mg_cycle=1))
for spec in ("6,6", "4,6", "4,4"):
    run("mlW2rel " + spec, {}, **ML(spec))
    run("mlW3rel " + spec, {}, **ML(spec, mg_wdepth=3))
run("mlF 6,6 cs1", {}, **ML("6,6", mg_cycle=1, mg_coarse_sweeps=1))
run("mlF 6,6 r128", {}, **ML("6,6", 127, mg_cycle=1))
