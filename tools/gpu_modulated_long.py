"""Where does the long modulated guide of config 5 stop converging?  Single-GPU solve(d::ModulatedDevice) and the 4-slab (thread
transport) solve of the notebook's Example-3 guide stretched to Nx x Ny cells (dh = 0.01), 3 sidebands.
    python tools/gpu_modulated_long.py 1024 512 2048 512 4096 512"""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import slab

dh, a, q = 0.01, 0.2202, 2.9263
w, Om = 2 * math.pi * 1.939e14, 4.541e14
NS = int(os.environ.get("NSLABS", 4))
args = [int(v) for v in sys.argv[1:]]
for Nx, Ny in zip(args[0::2], args[1::2]):
    g = fdfd.Grid(dh, [15, 15], [0.0, Nx * dh], [-Ny * dh / 2, Ny * dh / 2])
    d = fdfd.ModulatedDevice(g, w, Om, 1)
    xs = fdfd.xc(g)[:, None]; ys = fdfd.yc(g)[None, :]
    Lx = Nx * dh
    d.eps_r = np.where((ys >= -a / 2) & (ys <= a / 2), 12.25, 1.0) * np.ones((Nx, 1)) + 0j
    mod = (xs >= 0.1 * Lx) & (xs <= 0.9 * Lx) & (ys >= -a / 2) & (ys <= 0)
    d.deps_r = np.where(mod, np.exp(1j * q * xs) * np.ones((1, Ny)), 0)
    d.src = np.zeros((Nx, Ny), dtype=complex); d.src[25, :] = np.where(np.abs(ys[0]) <= 2 * a, 1j, 0)
    for tag, run in (("single GPU", lambda: fdfd.solve(d, maxit=int(os.environ.get("MAXIT", 6000)))[0][1].info),
                     (f"{NS} slabs (threads)", lambda: slab.solve_modulated_slabs_threads(d, NS, maxit=int(os.environ.get("MAXIT", 6000)))[1][0])):
        t0 = time.time()
        try:
            i = run()
            print(f"{Nx}x{Ny}x3 {tag}: flag={i['flag']} iters={i['iters']} relres={i['relres']:.2e} restarts={i['restarts']} krylov={i['solve_ms']:.0f} ms wall={time.time()-t0:.1f}s", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{Nx}x{Ny}x3 {tag}: {str(e)[-160:]} wall={time.time()-t0:.1f}s", flush=True)
