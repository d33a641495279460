"""Round 2 GPU experiment 5: single-launch line relaxation (corner rendezvous), hierarchy depth with F cycles."""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import _lib, workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = wl.synthetic_tm_device(fdfd, n, n, density=1 / 160.)
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
              f"ms/it={i['solve_ms']/max(1,i['iters']):.2f} us/launch={1e3*i['solve_ms']/max(1,i['launches']):.2f} {cyc}", flush=True)
        p.close()
    except Exception as e:
        print(f"n={n} {tag} {env} {kw}: FAILED {str(e)[-200:]}", flush=True)
    for k in (env or {}): del os.environ[k]
ML = dict(solver=_lib.SOLVER_MLKRYLOV, ml_spec=pack("6,6", 96))
run("bicg wd3", {}, solver=_lib.SOLVER_BICGSTAB)
run("ml F66", {}, **ML)
run("ml F66", {"FDFD_MG_KHSTOP": "2"}, **ML)
run("ml F66 cs4", {"FDFD_MG_KHSTOP": "2"}, mg_coarse_sweeps=4, **ML)
run("ml F66 cs1", {}, mg_coarse_sweeps=1, **ML)
run("ml W2rel", {"FDFD_ML_CYCLE": "2"}, mg_wdepth=2, **ML)
run("ml W2rel", {"FDFD_ML_CYCLE": "2", "FDFD_MG_KHSTOP": "2"}, mg_wdepth=2, **ML)
run("ml F88", {}, solver=_lib.SOLVER_MLKRYLOV, ml_spec=pack("8,8", 96))
run("ml F86", {}, solver=_lib.SOLVER_MLKRYLOV, ml_spec=pack("8,6", 96))
run("ml F64", {}, solver=_lib.SOLVER_MLKRYLOV, ml_spec=pack("6,4", 96))
run("bicg wd3", {"FDFD_MG_KHSTOP": "2"}, solver=_lib.SOLVER_BICGSTAB)
