"""multigrid as a stand-alone solver of the shifted system M u = f: per-cycle residual reduction (diagnostic)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
w = 2 * math.pi * 200e12
EPS0 = fdfd.EPS0
def conv(tag, g, eps, ordering=0, ncyc=8, beta=0.5, **kw):
    P = fdfd.Problem(g, fdfd.TM, w, eps, ordering=ordering, mg_beta=beta, **kw)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(g.N) + 1j * rng.standard_normal(g.N)
    def M(u):
        return fdfd.apply_operator(g, fdfd.TM, w, eps, u, ordering) - 1j * beta * w * w * EPS0 * g.L0 * eps * u
    f = M(x); u = np.zeros_like(x); h = []
    r = f.copy()
    for k in range(ncyc):
        u = u + P.precond(r)
        r = f - M(u)
        h.append(np.linalg.norm(r) / np.linalg.norm(f))
    fac = [h[0]] + [h[i] / h[i - 1] for i in range(1, len(h))]
    print(tag, kw, "factors:", " ".join("%.2f" % v for v in fac), " final %.1e" % h[-1], flush=True)
    # where does the residual sit?
    a = np.abs(r)
    print("    resid mean: interior %.2e  xstrip %.2e ystrip %.2e corner %.2e" % (a[30:-30, 30:-30].mean(), a[:16, 30:-30].mean(), a[30:-30, :16].mean(), a[:16, :16].mean()), flush=True)
    P.close()
for npml in ([15, 10], [15, 15], [10, 15], [8, 8], [15, 0], [0, 15]):
    g = fdfd.Grid(0.01, npml, [0.0, 4.0], [-1.0, 1.0])
    eps = np.ones(g.N, complex)
    conv(f"vac 400x200 npml={npml}", g, eps, mg_cycle=0)
g = fdfd.Grid(0.01, [15, 15], [0.0, 4.0], [-1.0, 1.0])
eps = np.ones(g.N, complex); eps[:, np.abs(fdfd.yc(g)) <= 0.11] = 12.25
for cyc in (0, 2):
    conv("wg 400x200 npml=[15,15]", g, eps, mg_cycle=cyc)
    conv("wg 400x200 npml=[15,15] f64", g, eps, mg_cycle=cyc, mg_precision=1)
conv("wg 400x200 npml=[15,15] bf", g, eps, ordering=1, mg_cycle=0)
g = fdfd.Grid(0.02, [15, 15], [0.0, 8.0], [-2.0, 2.0])
eps = np.ones(g.N, complex); eps[:, np.abs(fdfd.yc(g)) <= 0.11] = 12.25
conv("wg 400x200 dh.02", g, eps, mg_cycle=0)
conv("wg 400x200 dh.02 wl=1.0", g, eps, mg_cycle=0, mg_wline=1.0)
conv("wg 400x200 dh.02 wl=0.5", g, eps, mg_cycle=0, mg_wline=0.5)
