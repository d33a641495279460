"""parameter sweep on the bench workload: 4 frequencies solved concurrently (the bench step), wall time + iterations"""
import sys, os, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import workloads as wl
n = int(sys.argv[1])
d = wl.synthetic_tm_device(fdfd, n, n, density=1/160.)
d.omega = [2 * math.pi * (200e12 + 0.5e12 * k) for k in range(4)]
combos = [{}, {"mg_wdepth": 3}, {"mg_beta": 0.6}, {"mg_wdepth": 3, "mg_beta": 0.6}, {"mg_wdepth": 3, "mg_wjac": 0.7}, {"mg_beta": 0.6, "mg_wjac": 0.7},
          {"mg_wdepth": 3, "mg_beta": 0.6, "mg_wjac": 0.7}, {"mg_beta": 0.7}, {"mg_wdepth": 3, "mg_beta": 0.7}, {"mg_wdepth": 4}, {"mg_cycle": 1, "mg_beta": 0.6},
          {"mg_wdepth": 3, "mg_coarse_sweeps": 4}]
if len(sys.argv) > 2: combos = [eval(a) for a in sys.argv[2:]]
for kw in combos:
    t0 = time.time()
    try:
        fs = fdfd.solve(d, fdfd.TM, maxit=6000, concurrency=4, **kw)
        its = [f.info["iters"] for f in fs]; ok = all(f.info["flag"] == 0 for f in fs)
    except Exception as e:
        its, ok = str(e)[-60:], False
    print(f"n={n} {kw}: wall={time.time()-t0:.1f}s ok={ok} iters={its}", flush=True)
