python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -6
python tools/gpu_mgconv.py 2>&1 | grep -v resid | head -3
cat > /tmp/t7.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import fdfd_jl_b200 as fdfd
from importlib import import_module
wl = import_module("fdfd_jl_b200.workloads")
def run(tag, d, **cfg):
    P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, maxit=6000, **cfg); P.set_source(d.src); i = P.solve(); P.close()
    print(f"{tag} {cfg}: iters={i['iters']} relres={i['relres']:.1e} flag={i['flag']} ms={i['solve_ms']:.0f} ms/it={i['solve_ms']/max(1,i['iters']):.2f} launches/it={i['launches']/max(1,i['iters']):.0f} levels={i['mg_levels']} restarts={i['restarts']}", flush=True)
for n in (int(a) for a in sys.argv[1:]):
    d = wl.synthetic_tm_device(fdfd, n, n, density=1/160.)
    run(f"n{n}", d)
PY
python /tmp/t7.py 1024 2048 4096
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 800 --csv --log-file gpurun_out/launches_r01b.csv python tools/prof_solve.py 4096 4 > gpurun_out/prof_solve.log 2>&1
