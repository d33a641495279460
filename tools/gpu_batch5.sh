python tools/gpu_mgconv2.py 2>&1 | grep -v "coarse=300"
python tools/gpu_mgconv.py 2>&1 | grep -v resid
python tools/gpu_diag4.py 2>&1
python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -8
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import fdfd_jl_b200 as fdfd
from importlib import import_module
wl = import_module("fdfd_jl_b200.workloads")
for n, dens in ((1024, 1/40.), (1024, 1/160.), (2048, 1/160.)):
    d = wl.synthetic_tm_device(fdfd, n, n, density=dens)
    for cfg in (dict(mg_cycle=0), dict(mg_cycle=2, mg_wdepth=2), dict(mg_cycle=2, mg_wdepth=3), dict(mg_cycle=1)):
        P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, maxit=4000, **cfg); P.set_source(d.src); i = P.solve(); P.close()
        print(f"synth n={n} d={dens:.4f} {cfg}: iters={i['iters']} relres={i['relres']:.1e} flag={i['flag']} ms={i['solve_ms']:.0f} ms/it={i['solve_ms']/max(1,i['iters']):.2f} launches/it={i['launches']/max(1,i['iters']):.0f}", flush=True)
PY
