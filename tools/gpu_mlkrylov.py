"""GPU run for round 2: the multilevel Krylov solver (FDFD_SOLVER_MLKRYLOV, csrc/mlkrylov.cu) against the default
BiCGSTAB + multigrid on the bench workload -- iterations, solve time, multigrid cycles per level.

    gpurun --timeout 900 -- 'python tools/gpu_mlkrylov.py 1024 2048 4096 > gpurun_out/mlkrylov.log 2>&1'
    python tools/gpu_mlkrylov.py 4096 --spec 6,12 8,16 6,6,12 --restart 96        # spec sweep at one size
"""
import argparse, ctypes as C, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import _lib, workloads as wl

ap = argparse.ArgumentParser()
ap.add_argument("sizes", type=int, nargs="+")
ap.add_argument("--spec", nargs="*", default=["6,12"], help="FGMRES steps on levels 1,2[,3], e.g. 6,12")
ap.add_argument("--restart", type=int, default=96)
ap.add_argument("--no-baseline", action="store_true")
ap.add_argument("--maxit", type=int, default=6000)
a = ap.parse_args()


def pack(spec, restart):
    k = [int(x) for x in spec.split(",")] + [0, 0, 0]
    return k[0] | (k[1] << 8) | (k[2] << 16) | (restart << 24)


for n in a.sizes:
    d = wl.synthetic_tm_device(fdfd, n, n, density=1 / 160.)
    ref = None
    if not a.no_baseline:
        t0 = time.time()
        ref = fdfd.solve(d, fdfd.TM, maxit=a.maxit)
        i = ref.info
        print(f"n={n} BiCGSTAB+MG : flag={i['flag']} iters={i['iters']} relres={i['relres']:.2e} solve={i['solve_ms']:.0f} ms "
              f"setup={i['setup_ms']:.0f} ms wall={time.time() - t0:.1f}s launches={i['launches']}", flush=True)
    for spec in a.spec:
        t0 = time.time()
        try:
            # Problem handle so that the cycle counters can be read back
            p = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, solver=_lib.SOLVER_MLKRYLOV, ml_spec=pack(spec, a.restart), maxit=a.maxit)
            p.set_source(d.src)
            i = p.solve()
            cyc = p.ml_cycles()
            x = p.solution()
            err = float(np.linalg.norm(x - ref.data[:, :, 0]) / np.linalg.norm(ref.data[:, :, 0])) if ref is not None else float("nan")
            print(f"n={n} MLKRYLOV {spec:>8s} restart {a.restart}: flag={i['flag']} outer={i['iters']} relres={i['relres']:.2e} "
                  f"solve={i['solve_ms']:.0f} ms restarts={i['restarts']} cycles/level={cyc} launches={i['launches']} "
                  f"wall={time.time() - t0:.1f}s |x-x_ref|/|x_ref|={err:.1e}", flush=True)
            p.close()
        except Exception as e:  # noqa: BLE001
            print(f"n={n} MLKRYLOV {spec}: FAILED {str(e)[-200:]}", flush=True)
