import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from importlib import import_module
wl = import_module("fdfd_jl_b200.workloads")
def log(*a): print(*a, flush=True)
out = {}
def run(tag, d, **cfg):
    try:
        P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, **cfg)
        P.set_source(d.src)
        info = P.solve()
        log(f"{tag} {cfg}: iters={info['iters']} relres={info['relres']:.2e} flag={info['flag']} solve_ms={info['solve_ms']:.1f} "
            f"ms/it={info['solve_ms']/max(1,info['iters']):.3f} launches/it={info['launches']/max(1,info['iters']):.0f} restarts={info['restarts']} levels={info['mg_levels']}")
        out[tag + json.dumps(cfg)] = info
        P.close()
    except Exception as e:
        log(tag, cfg, "EXC", e)

d1 = wl.synthetic_tm_device(fdfd, 1024, 1024)
base = dict(maxit=2500)
# graph on/off
run("n1024", d1, mg_cycle=0, use_graph=0, **base)
run("n1024", d1, mg_cycle=0, use_graph=1, **base)
for cyc, wd in ((0, 0), (1, 0), (2, 2)):
    for growth in (0.0, 0.125, 0.25, 0.5, 1.0):
        run("n1024", d1, mg_cycle=cyc, mg_wdepth=wd, mg_shift_growth=growth, **base)
for wj in (0.5, 0.65):
    run("n1024", d1, mg_cycle=0, mg_wjac=wj, **base)
    run("n1024", d1, mg_cycle=0, mg_wjac=wj, mg_nu1=2, mg_nu2=2, **base)
for growth in (0.25, 0.5):
    run("n1024", d1, mg_cycle=0, mg_shift_growth=growth, mg_nu1=2, mg_nu2=2, **base)
    for ml in (5, 6, 7):
        run("n1024", d1, mg_cycle=0, mg_shift_growth=growth, mg_max_levels=ml, mg_coarse_sweeps=8, **base)
for beta in (0.7, 1.0):
    run("n1024", d1, mg_cycle=0, mg_beta=beta, **base)
d1s = wl.synthetic_tm_device(fdfd, 1024, 1024, density=1.0 / 160)
for cyc, wd in ((0, 0), (1, 0), (2, 2)):
    for growth in (0.0, 0.25):
        run("n1024sparse", d1s, mg_cycle=cyc, mg_wdepth=wd, mg_shift_growth=growth, **base)
json.dump(out, open("gpurun_out/explore2.json", "w"), indent=1)
