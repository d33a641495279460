import sys, time, json, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
def log(*a): print(*a, flush=True)
w, Om = 2 * math.pi * 1.939e14, 4.541e14
a, q = 0.2202, 2.9263
def dev(ns, coupled=True, L=15.0, dh=0.01, npml=(15, 10)):
    g = fdfd.Grid(dh, list(npml), [0.0, L], [-1.0, 1.0])
    d = fdfd.ModulatedDevice(g, w, Om, ns)
    fdfd.setup_eps_r(d, lambda x, y: -a / 2 <= y <= a / 2, 12.25)
    if coupled:
        fdfd.setup_deps_r(d, lambda x, y: (0.1 * L <= x <= 0.78 * L) and (-a / 2 <= y <= 0), lambda x, y: np.exp(1j * q * x))
    fdfd.add_mode(d, fdfd.Mode(fdfd.TM, fdfd.XHAT, 3.5, fdfd.Point(0.2, 0), 4 * a))
    return d
def run(tag, d, **kw):
    t0 = time.time()
    try:
        f = fdfd.solve(d, maxit=kw.pop("maxit", 3000), **kw)[0]
        i = f[0].info
        log(f"{tag} {kw}: iters={i['iters']} relres={i['relres']:.2e} flag={i['flag']} ms={i['solve_ms']:.0f} levels={i['mg_levels']} t={time.time()-t0:.1f}")
    except Exception as e:
        log(f"{tag} {kw}: EXC {str(e)[-120:]} t={time.time()-t0:.1f}")
run("ns0 15um dh.01", dev(0))
run("ns0 15um dh.01 f64", dev(0), mg_precision=1)
run("ns0 15um dh.01 W2", dev(0), mg_wdepth=2)
run("ns0 15um dh.01 F", dev(0), mg_cycle=1)
run("ns0 15um dh.02", dev(0, dh=0.02))
run("ns0 5um dh.01", dev(0, L=5.0))
run("ns0 15um dh.01 npml15,15", dev(0, npml=(15, 15)))
run("ns1 uncoupled 15um", dev(1, coupled=False))
run("ns1 coupled 15um", dev(1))
run("ns1 coupled 15um dh.02", dev(1, dh=0.02))
run("ns1 coupled 5um dh.01", dev(1, L=5.0))
