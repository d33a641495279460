"""Round 2 GPU experiment 2: hierarchy depth / W-depth / coarse sweeps against wall time (the coarse levels are latency-bound:
~5 us per launch, ~280 small launches per BiCGSTAB iteration), BiCGSTAB and multilevel Krylov, single stream."""
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
        print(f"n={n} {tag} {env or ''} {kw}: flag={i['flag']} iters={i['iters']} solve={i['solve_ms']:.0f} ms levels={i['mg_levels']} launches={i['launches']} "
              f"ms/it={i['solve_ms']/max(1,i['iters']):.2f}", flush=True)
        p.close()
    except Exception as e:
        print(f"n={n} {tag} {env} {kw}: FAILED {str(e)[-200:]}", flush=True)
    for k in (env or {}): del os.environ[k]

ML = dict(solver=_lib.SOLVER_MLKRYLOV, ml_spec=pack("6,6", 96))
for env in ({}, {"FDFD_MG_KHSTOP": "2"}, {"FDFD_MG_KHSTOP": "1"}, {"FDFD_MG_KHSTOP": "0.5"}):
    run("bicg", env)
    run("ml66", env, **ML)
for kw in (dict(mg_wdepth=1), dict(mg_wdepth=1, mg_coarse_sweeps=4), dict(mg_coarse_sweeps=1), dict(mg_wdepth=3), dict(mg_cycle=0), dict(mg_cycle=1)):
    run("bicg", {}, **kw)
    run("ml66", {}, **dict(ML, **kw))
run("bicg", {"FDFD_MG_KHSTOP": "2"}, mg_coarse_sweeps=4)
run("ml66", {"FDFD_MG_KHSTOP": "2"}, mg_coarse_sweeps=4, **ML)
