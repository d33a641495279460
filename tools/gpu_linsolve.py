"""GPU run of the dolinearsolve seam (csrc/linsolve.cu, written blind in round 1): for the bench map at the given grid edges
  * k_sell_spmv: ms per launch and GB/s against the measured HBM peak (the FDFD TM matrix, 5 entries per row, 1-based CSC as Julia
    hands it over), beside the matrix-free stencil on the same operator,
  * the grid-hinted seam (multigrid path) against fdfd.solve on the same device: iterations, ms, agreement of the solutions,
  * the generic Jacobi path on the smallest size only (thousands of iterations on a Helmholtz matrix).
    gpurun --timeout 900 -- 'python tools/gpu_linsolve.py 512 1024 2048'
"""
import ctypes as C
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fdfd_jl_b200 as fdfd  # noqa: E402
from importlib import import_module  # noqa: E402

wl = import_module("fdfd_jl_b200.workloads")


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [512, 1024]
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    ctx = fdfd.default_context()
    L = fdfd.lib()
    for k, n in enumerate(sizes):
        d = wl.synthetic_tm_device(fdfd, n, n, density=1.0 / 160.0)   # the bench map
        g, w = d.grid, d.omega[0]
        N = n * n
        colptr, rowval, nzval = fdfd.assemble_system(g, fdfd.TM, w, d.eps_r, fmt=fdfd._lib.CSC, index_base=1)
        b = 1j * w * np.asarray(d.src).ravel(order="F")
        ms, ab = C.c_double(0), C.c_double(0)
        fdfd.check(L.fdfd_debug_sell_bench(ctx.handle, N, fdfd.ptr(colptr), fdfd.ptr(rowval), fdfd.ptr(nzval), 1, 100, C.byref(ms), C.byref(ab)),
                   ctx.handle)
        P = fdfd.Problem(g, fdfd.TM, w, d.eps_r, ctx=ctx, precond=0)
        ms_st = P.bench_apply(100)
        P.close()
        print(f"[{n}^2] k_sell_spmv {ms.value:.4f} ms/launch = {ab.value / ms.value / 1e6:.0f} GB/s = {ab.value / ms.value / 1e6 / peak:.2f} of {peak:.0f}"
              f" ({ab.value / N:.0f} B/row);  matrix-free k_apply {ms_st:.4f} ms = {48.0 * N / ms_st / 1e6:.0f} GB/s (48 B/pt)", flush=True)
        t0 = time.perf_counter()
        x, info = fdfd.dolinearsolve((colptr, rowval, nzval), b, index_base=1, grid=g, omega=w, return_info=True)
        t1 = time.perf_counter()
        f = fdfd.solve(d, fdfd.TM)
        t2 = time.perf_counter()
        ez = f["Ez"].ravel(order="F")
        print(f"[{n}^2] grid-hinted seam: mg_levels {info['mg_levels']} iters {info['iters']} relres(A) {info['relres']:.2e} "
              f"krylov {info['solve_ms']:.0f} ms, call {1e3 * (t1 - t0):.0f} ms | solve(d): iters {f.info['iters']} krylov {f.info['solve_ms']:.0f} ms, "
              f"call {1e3 * (t2 - t1):.0f} ms | |x - Ez|/|Ez| = {np.linalg.norm(x - ez) / np.linalg.norm(ez):.2e}", flush=True)
        if k == 0 and n <= 512:
            try:
                t0 = time.perf_counter()
                xj, ij = fdfd.dolinearsolve((colptr, rowval, nzval), b, index_base=1, maxit=200000, check_every=64, return_info=True)
                print(f"[{n}^2] generic Jacobi path: iters {ij['iters']} relres {ij['relres']:.2e} krylov {ij['solve_ms']:.0f} ms "
                      f"({ij['solve_ms'] / max(1, ij['iters']):.3f} ms/it), call {1e3 * (time.perf_counter() - t0):.0f} ms, "
                      f"|x - Ez|/|Ez| = {np.linalg.norm(xj - ez) / np.linalg.norm(ez):.2e}", flush=True)
            except fdfd.FdfdError as e:
                print(f"[{n}^2] generic Jacobi path: {e}", flush=True)


if __name__ == "__main__":
    main()
