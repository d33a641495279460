"""driver for ncu: a few un-graphed BiCGSTAB iterations at 4096^2 (the first multigrid launches of a solve are the level-0 kernels)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdfd_jl_b200 as fdfd
from importlib import import_module
wl = import_module("fdfd_jl_b200.workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
its = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d = wl.synthetic_tm_device(fdfd, n, n, density=1/160)
P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, maxit=its, use_graph=0, check_every=its, solver=fdfd._lib.SOLVER_BICGSTAB)
P.set_source(d.src)
print(P.solve())
