import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import workloads as wl
n = int(sys.argv[1])
d = wl.synthetic_tm_device(fdfd, n, n, density=1/160.)
for kw in ({}, {"mg_wdepth": 1}, {"mg_wdepth": 3}, {"mg_cycle": 1}, {"mg_cycle": 0, "mg_nu1": 2, "mg_nu2": 2}, {"mg_wdepth": 1, "mg_nu2": 2}, {"mg_beta": 0.4}, {"mg_beta": 0.6}, {"mg_wjac": 0.7}, {"mg_wjac": 0.9}):
    P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, maxit=8000, **kw); P.set_source(d.src); i = P.solve(); P.close()
    print(f"n={n} {kw}: iters={i['iters']} flag={i['flag']} ms={i['solve_ms']:.0f} ms/it={i['solve_ms']/max(1,i['iters']):.2f} launches/it={i['launches']/max(1,i['iters']):.0f}", flush=True)
