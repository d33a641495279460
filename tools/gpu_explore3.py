import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdfd_jl_b200 as fdfd
from importlib import import_module
wl = import_module("fdfd_jl_b200.workloads")
def log(*a): print(*a, flush=True)
def run(tag, d, **cfg):
    try:
        P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, **cfg)
        P.set_source(d.src)
        info = P.solve()
        log(f"{tag} {cfg}: iters={info['iters']} relres={info['relres']:.2e} flag={info['flag']} solve_ms={info['solve_ms']:.1f} "
            f"ms/it={info['solve_ms']/max(1,info['iters']):.3f} launches/it={info['launches']/max(1,info['iters']):.0f} setup_ms={info['setup_ms']:.0f} restarts={info['restarts']} levels={info['mg_levels']}")
        P.close()
    except Exception as e:
        log(tag, cfg, "EXC", e)
for dens in (1/160., 1/40.):
    d = wl.synthetic_tm_device(fdfd, 2048, 2048, density=dens)
    run(f"n2048 d={dens:.4f}", d, mg_cycle=2, mg_wdepth=2, maxit=6000)
    run(f"n2048 d={dens:.4f}", d, mg_cycle=2, mg_wdepth=3, maxit=6000)
d = wl.synthetic_tm_device(fdfd, 4096, 4096, density=1/160.)
run("n4096 d=1/160", d, mg_cycle=2, mg_wdepth=2, maxit=15000)
run("n4096 d=1/160", d, mg_cycle=0, maxit=15000)
