import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdfd_jl_b200 as fdfd
from importlib import import_module
wl = import_module("fdfd_jl_b200.workloads")
for n in (2048, 4096):
    d = wl.synthetic_tm_device(fdfd, n, n, density=1/160.)
    for cfg in (dict(), dict(mg_wdepth=3), dict(mg_wdepth=4), dict(mg_cycle=1), dict(mg_wdepth=1)):
        P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, maxit=8000, **cfg); P.set_source(d.src); i = P.solve(); P.close()
        print("n=%d %s: iters=%d ms=%.0f ms/it=%.2f launches/it=%.0f flag=%d" % (n, cfg, i["iters"], i["solve_ms"], i["solve_ms"] / i["iters"], i["launches"] / i["iters"], i["flag"]), flush=True)
