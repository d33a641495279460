cat > /tmp/t9.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import fdfd_jl_b200 as fdfd
from importlib import import_module
wl = import_module("fdfd_jl_b200.workloads")
n = int(sys.argv[1])
d = wl.synthetic_tm_device(fdfd, n, n, density=1/160.)
for cs in (1, 2, 3, 4):
    P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, maxit=6000, mg_coarse_sweeps=cs); P.set_source(d.src); i = P.solve(); P.close()
    print(f"n={n} coarse_sweeps={cs} env={os.environ.get('FDFD_MG_PAD','-')},{os.environ.get('FDFD_MG_KHSTOP','-')}: iters={i['iters']} ms={i['solve_ms']:.0f} ms/it={i['solve_ms']/i['iters']:.2f} launches/it={i['launches']/i['iters']:.0f} levels={i['mg_levels']}", flush=True)
PY
python /tmp/t9.py 2048
FDFD_MG_PAD=0 python /tmp/t9.py 2048
FDFD_MG_KHSTOP=3 python /tmp/t9.py 2048
FDFD_MG_KHSTOP=6 python /tmp/t9.py 2048
