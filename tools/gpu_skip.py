"""timing diagnostics: fixed number of BiCGSTAB iterations of B concurrent 4096^2 solves; run with FDFD_MG_SKIP=..."""
import sys, os, time, math, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
its = int(sys.argv[3]) if len(sys.argv) > 3 else 240
d = wl.synthetic_tm_device(fdfd, n, n, density=1 / 160)
ctx = fdfd.default_context()
g = d.grid; gc = g.as_c(); N = n * n
om = (C.c_double * B)(*[2 * math.pi * (200e12 + 0.5e12 * k) for k in range(B)])
opts = fdfd.default_opts(concurrency=B, maxit=its, check_every=its)
import torch
eps = torch.from_numpy(np.asfortranarray(d.eps_r).ravel(order="F").copy()).cuda()
src = torch.from_numpy(np.asfortranarray(d.src).ravel(order="F").copy()).cuda()
out = torch.empty(B * 3 * N, dtype=torch.complex128, device="cuda")
infos = (fdfd.Info * B)()
for rep in range(2):
    t0 = time.time()
    fdfd.lib().fdfd_solve_driven(ctx.handle, C.byref(gc), fdfd.TM, B, om, fdfd.ptr(eps.data_ptr()), fdfd.ptr(src.data_ptr()), 0, C.byref(opts), fdfd.ptr(out.data_ptr()), infos)
    torch.cuda.synchronize()
    t1 = time.time()
print(f"skip={os.environ.get('FDFD_MG_SKIP','0')} n={n} B={B} its={its}: krylov_ms={[round(i.solve_ms) for i in infos]} -> {max(i.solve_ms for i in infos)/its:.3f} ms per iteration-round, {max(i.solve_ms for i in infos)/its/B:.3f} ms per iteration-solve; wall {t1-t0:.1f}s", flush=True)
