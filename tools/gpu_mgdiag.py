"""GPU diagnostic (round 2): the CUDA multigrid cycle against the NumPy prototype cycle on the same vectors.
Why: at 512^2 the prototype's BiCGSTAB needs 92 iterations, the GPU path 161-167 -- which part of the cycle differs?
    gpurun -- 'python tools/gpu_mgdiag.py 512 > gpurun_out/mgdiag.log 2>&1'
"""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import _lib, workloads as wl
from oracle import fdfd_oracle as O
from tools.mg_prototype import MG, Level, synth_device

n = int(sys.argv[1])
d = wl.synthetic_tm_device(fdfd, n, n, density=1 / 160.)
do = synth_device(n, n); do.eps_r[:] = d.eps_r
g = do.grid; omega = do.omega[0]
eps0, mu0, _ = O.normalize_parameters(g)
cxm, cxp, cym, cyp = O.stencil_coefficients(g, omega, "fb")
Mop = Level(n, n, cxm, cxp, cym, cyp, (1 - 0.5j) * omega ** 2 * eps0 * do.eps_r)
b = 1j * omega * do.src
mgp = MG(g, omega, do.eps_r, beta=0.5, wj=0.7, wl=0.6, nu1=1, nu2=1, cycle="W", wdepth=2, coarse_sweeps=2, min_n=8, dtype=np.complex128)
rng = np.random.default_rng(1)
vecs = {"b": b, "rand": rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)),
        "smooth": np.exp(1j * 4.0 * O.xc(g))[:, None] * np.cos(2.0 * O.yc(g))[None, :] + 0j}
npml = 15
def regions(r):
    m = np.zeros((n, n), bool); m[npml + 2:n - npml - 2, npml + 2:n - npml - 2] = True
    return np.linalg.norm(r[m]), np.linalg.norm(r[~m])

def study(tag, **kw):
    t0 = time.time()
    p = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, **kw)
    p.set_source(d.src)
    i = p.solve()
    print(f"[{tag}] solve: iters={i['iters']} relres={i['relres']:.2e} levels={i['mg_levels']} restarts={i['restarts']} ms={i['solve_ms']:.0f}", flush=True)
    for name, v in vecs.items():
        ug = p.precond(v); up = mgp(v)
        rg = v - Mop.apply(ug); rp = v - Mop.apply(up)
        nv = np.linalg.norm(v)
        print(f"[{tag}] {name:6s}: |v-M u|/|v| gpu {np.linalg.norm(rg)/nv:.3f} proto {np.linalg.norm(rp)/nv:.3f}  |ug-up|/|up| {np.linalg.norm(ug-up)/np.linalg.norm(up):.3f}"
              f"  gpu resid interior/pml {regions(rg)[0]/nv:.3f}/{regions(rg)[1]/nv:.3f} proto {regions(rp)[0]/nv:.3f}/{regions(rp)[1]/nv:.3f}", flush=True)
    # stationary iteration factor
    for name, cyc in (("gpu", p.precond), ("proto", mgp)):
        v = vecs["rand"]; u = np.zeros_like(v); r = v.copy(); hist = []
        for k in range(12):
            u = u + cyc(r); r = v - Mop.apply(u); hist.append(np.linalg.norm(r) / np.linalg.norm(v))
        print(f"[{tag}] stationary {name}: " + " ".join(f"{h:.2e}" for h in hist) + f"  last factor {hist[-1]/hist[-2]:.3f}", flush=True)
    p.close()

study("default")





study("nu2", mg_nu1=2, mg_nu2=2)
