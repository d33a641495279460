import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
w = 2 * math.pi * 200e12
EPS0 = fdfd.EPS0
def conv(tag, g, eps, ordering=0, ncyc=7, beta=0.5, **kw):
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
    P.close()
g = fdfd.Grid(0.01, [15, 10], [0.0, 4.0], [-1.0, 1.0])
eps = np.ones(g.N, complex)
for ml in (2, 3, 4, 5, 6, 7, 8):
    conv(f"vac 400x200 levels={ml} coarse=300", g, eps, mg_cycle=0, mg_max_levels=ml, mg_coarse_sweeps=300)
for ml in (4, 6, 8):
    conv(f"vac 400x200 levels={ml} coarse=4", g, eps, mg_cycle=0, mg_max_levels=ml, mg_coarse_sweeps=4)
conv("vac V(2,2)", g, eps, mg_cycle=0, mg_nu1=2, mg_nu2=2)
conv("vac wj=.5", g, eps, mg_cycle=0, mg_wjac=0.5)
conv("vac wl=.5", g, eps, mg_cycle=0, mg_wline=0.5)
conv("vac wj=.5 wl=.5", g, eps, mg_cycle=0, mg_wjac=0.5, mg_wline=0.5)
conv("vac F", g, eps, mg_cycle=1)
conv("vac W1", g, eps, mg_cycle=2, mg_wdepth=1)
conv("vac W2", g, eps, mg_cycle=2, mg_wdepth=2)
g2 = fdfd.Grid(0.01, [15, 10], [0.0, 5.12], [-1.28, 1.28])  # 512 x 256: no odd sizes until 4x2
conv("vac 512x256", g2, np.ones(g2.N, complex), mg_cycle=0)
g3 = fdfd.Grid(0.0234375, [15, 15], [-3, 3], [-3, 3])  # the prototype's 256^2 case
conv("vac 256^2 proto", g3, np.ones(g3.N, complex), mg_cycle=0)
