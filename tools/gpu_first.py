"""first GPU exploration run: apply bandwidth + solver convergence/timing on synthetic devices."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from importlib import import_module
wl = import_module("fdfd_jl_b200.workloads")

out = {}
def log(*a):
    print(*a, flush=True)

for n in (1024, 2048, 4096):
    d = wl.synthetic_tm_device(fdfd, n, n)
    P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, precond=0)
    ms = P.bench_apply(50)
    gbs = 48.0 * n * n / (ms * 1e-3) / 1e9
    log(f"apply {n}^2: {ms:.4f} ms  {gbs:.0f} GB/s")
    out[f"apply_{n}"] = dict(ms=ms, gbs=gbs)
    P.close()

configs = [dict(mg_cycle=0), dict(mg_cycle=1), dict(mg_cycle=2, mg_wdepth=2), dict(mg_cycle=2, mg_wdepth=4),
           dict(mg_cycle=0, mg_precision=1), dict(mg_cycle=0, mg_nu1=2, mg_nu2=2)]
for n in (512, 1024, 2048):
    d = wl.synthetic_tm_device(fdfd, n, n)
    for cfg in configs:
        t0 = time.time()
        P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, maxit=6000, **cfg)
        P.set_source(d.src)
        info = P.solve()
        log(f"solve {n}^2 {cfg}: iters={info['iters']} relres={info['relres']:.2e} flag={info['flag']} solve_ms={info['solve_ms']:.1f} "
            f"ms/it={info['solve_ms']/max(1,info['iters']):.3f} setup_ms={info['setup_ms']:.1f} launches={info['launches']} restarts={info['restarts']} levels={info['mg_levels']}")
        out[f"solve_{n}_{json.dumps(cfg)}"] = info
        P.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/first.json", "w"), indent=1)
