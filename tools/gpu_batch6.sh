python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import fdfd_jl_b200 as fdfd
from importlib import import_module
wl = import_module("fdfd_jl_b200.workloads")
def run(tag, d, **cfg):
    P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, maxit=6000, **cfg); P.set_source(d.src); i = P.solve(); P.close()
    print(f"{tag} {cfg}: iters={i['iters']} relres={i['relres']:.1e} flag={i['flag']} ms={i['solve_ms']:.0f} ms/it={i['solve_ms']/max(1,i['iters']):.2f} launches/it={i['launches']/max(1,i['iters']):.0f} levels={i['mg_levels']} restarts={i['restarts']}", flush=True)
d = wl.synthetic_tm_device(fdfd, 2048, 2048, density=1/160.)
for ml in (5, 6, 7, 8):
    for cs in (4, 8):
        run("n2048 trunc", d, mg_max_levels=ml, mg_coarse_sweeps=cs)
run("n2048 full", d)
run("n2048 nu2", d, mg_nu1=2, mg_nu2=2)
run("n2048 nu21", d, mg_nu1=2, mg_nu2=1)
run("n2048 beta.4", d, mg_beta=0.4)
run("n2048 beta.6", d, mg_beta=0.6)
run("n2048 wj.9", d, mg_wjac=0.9)
run("n2048 wj.7", d, mg_wjac=0.7)
d = wl.synthetic_tm_device(fdfd, 4096, 4096, density=1/160.)
run("n4096 full", d)
run("n4096 trunc7", d, mg_max_levels=7, mg_coarse_sweeps=6)
run("n4096 W3", d, mg_wdepth=3)
PY
python bench.py --steps 2 --warmup 1 2>&1 | tail -3
