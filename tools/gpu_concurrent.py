import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from importlib import import_module
wl = import_module("fdfd_jl_b200.workloads")
n = int(sys.argv[1])
d = wl.synthetic_tm_device(fdfd, n, n, density=1/160.)
import math
def work(k, out, wd):
    ctx = fdfd.Context(0)
    w = 2 * math.pi * (200e12 + 0.5e12 * k)
    P = fdfd.Problem(d.grid, fdfd.TM, w, d.eps_r, ctx=ctx, mg_wdepth=wd)
    P.set_source(d.src)
    out[k] = P.solve()
    P.close()
for wd in (2, 3):
    for B in (1, 2, 4):
        out = [None] * B
        th = [threading.Thread(target=work, args=(k, out, wd)) for k in range(B)]
        t0 = time.time()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.time() - t0
        print("n=%d wdepth=%d B=%d: wall %.2f s -> %.3f solves/s ; iters %s ; per-solve krylov ms %s" % (n, wd, B, dt, B / dt, [o["iters"] for o in out], ["%.0f" % o["solve_ms"] for o in out]), flush=True)
