#!/usr/bin/env python
"""bench.py -- solves/sec and stencil HBM GB/s of the 4096^2 TM solve to a 1e-10 residual (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference algorithm (sparse direct LU) on host cores

One "step" = one call of the hot path, `solve(device, TM)` == fdfd_solve_driven, on the synthetic 4096x4096 TM
device carrying a sweep of --sweep (default 4) frequencies: for each frequency the per-frequency operator
setup (PML coefficients, multigrid hierarchy), the GPU-resident Krylov solve to ||b-Ax||/||b|| <= 1e-10 (checked with
the fp64 operator) and H-field recovery; the library overlaps the frequencies on separate streams.  The metric counts
solved (omega, source) right-hand sides per second.  `value` times that with eps_r/src resident in HBM; `e2e` times the
public API call (`solve(device)`) with host buffers, H2D and D2H copies inside the timed region.  N > 1: one rank
per GPU, every rank solves a sweep of the same --sweep frequencies on its own copy of the device (independent units, no
data-path collective): weak scaling with identical work per GPU.  (Round 1 gave rank r the frequencies 200 + 0.5 (4 r + k) THz:
the iteration count grows with frequency, so the MAX over ranks measured the hardest frequency set, not the machine.)
With --slab-grid G (default 8192 when N > 1) the same launch also times ONE G x G solve split into row slabs over the N ranks
(halo exchange + allreduce over NCCL, strong scaling) and reports it under "slab" in the same JSON line.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "solves_per_sec_4096x4096_TM_to_1e-10"  # --grid other than 4096 renames the metric (debug runs only)
UNIT = "solves/s"
ALG_BYTES_PER_POINT = 48.0  # read x 16 + read w^2*eps 16 + write y 16 (complex128), SURVEY §8d


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", dest="n", type=int, default=4096, help="grid edge (cells); 4096 is the metric's configuration")
    ap.add_argument("--density", type=float, default=1.0 / 160.0, help="scatterers per um^2 of the synthetic map")
    ap.add_argument("--ref-grid", dest="ref_n", type=int, default=512, help="grid edge of the bounded CPU sample (one step of the reference arm)")
    ap.add_argument("--ref-fit", default="256,512,1024", help="reference arm: grid edges timed once each for the in-run scaling-law fit")
    ap.add_argument("--slab-grid", type=int, default=-1, help="also time one G^2 solve split into row slabs over the ranks (0: off; default: 8192 when --gpus > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", type=int, default=4, help="frequencies per GPU per step, solved concurrently (one stream each)")
    ap.add_argument("--no-e2e", action="store_true", help="slab mode: skip the host-buffer (end-to-end) repetition of the solve")
    ap.add_argument("--solver", default="auto", choices=["auto", "bicgstab", "mlkrylov"],
                    help="sweep mode: auto (library default: multilevel Krylov from 2^22 grid points, BiCGSTAB below), or force one")
    ap.add_argument("--ml-spec", default="", help="--solver mlkrylov: FGMRES steps on levels 1,2[,3] (default: library default 6,6)")
    ap.add_argument("--ml-restart", type=int, default=0, help="--solver mlkrylov: level-0 restart length (0: a 1/concurrency share of the free HBM, at most 96)")
    ap.add_argument("--maxit", type=int, default=0, help="Krylov iteration cap (0: library default 20000)")
    ap.add_argument("--concurrency", type=int, default=0, help="frequencies solved at the same time (default: --sweep)")
    ap.add_argument("--mode", default="sweep", choices=["sweep", "slab"],
                    help="sweep (default, the metric): disjoint frequencies per GPU, weak scaling.  slab: ONE --grid^2 solve split "
                         "into row slabs over the GPUs (halo exchange + allreduce over NCCL), strong scaling (BASELINE config 5)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------
def cpu_reference_solve(n, density, threads=None):
    """the reference algorithm on host cores: assemble A exactly as driven.jl:21-36 does, sparse direct LU (SuperLU
    standing in for Julia's UMFPACK `lu(A)\\b`: Julia/UMFPACK/Pardiso are not installable here), H recovery."""
    from oracle import fdfd_oracle as O
    import fdfd_jl_b200 as fdfd
    from importlib import import_module
    wl = import_module("fdfd_jl_b200.workloads")
    d = wl.synthetic_tm_device(fdfd, n, n, density=density)  # host-side numpy map (no GPU involved)
    go = O.Grid2D(0.02, [15, 15], [0.0, n * 0.02], [0.0, n * 0.02])
    do = O.Device(go, list(d.omega))
    do.eps_r[:] = d.eps_r
    do.src[:] = d.src
    t0 = time.perf_counter()
    f = O.solve(do, O.TM)
    dt = time.perf_counter() - t0
    return dt, f


def fit_exponent(samples):
    """least-squares slope of log t against log N (N = unknowns) over the measured (edge, seconds) samples"""
    if len(samples) < 2:
        return math.log(5.5) / math.log(4.0), "survey-time law (x5.5 per 4x unknowns, BASELINE.md §3): fewer than two sizes measured"
    x = np.log([float(n) * n for n, _ in samples]); y = np.log([t for _, t in samples])
    return float(np.polyfit(x, y, 1)[0]), "fitted in this run over " + ", ".join(f"{n}^2: {t:.2f} s" for n, t in samples)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    cores = os.cpu_count() or 1
    nref = args.ref_n
    times = []
    for i in range(args.warmup + args.steps):
        dt, _ = cpu_reference_solve(nref, args.density)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    # the direct solver's scaling law, measured once per size in this run on this box (not part of the K timed steps)
    samples = []
    for ne in sorted({int(v) for v in args.ref_fit.split(",") if v.strip()}):
        samples.append((ne, t) if ne == nref else (ne, cpu_reference_solve(ne, args.density)[0]))
    expo, how = fit_exponent(samples)
    scale = ((args.n * args.n) / float(nref * nref)) ** expo
    val = 1.0 / (t * scale)
    sample = (f"sparse direct LU (SciPy SuperLU standing in for Julia \\ / UMFPACK) of the same synthetic TM device, SOLVED at "
              f"{nref}x{nref}: {t:.2f} s per solve on {cores} host cores; extrapolated to {args.n}^2 by t ~ N^{expo:.3f} ({how}) "
              f"-> {t * scale:.0f} s per solve" + ("; the 4096^2 factorisation itself needs >200 GB of host RAM and ~10 h" if args.n >= 4096 else ""))
    cfg = workload_config(args)
    cfg["grid_solved_by_this_arm"] = [nref, nref]
    cfg["extrapolated_to"] = [args.n, args.n]
    line = {"impl": "reference", "metric": metric_name(args), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * scale * 1e3, "measured_ms_per_step_at_sample_grid": t * 1e3,
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "sample_seconds_per_solve": t, "sample_grid": [nref, nref], "fit_exponent": expo,
                             "fit_samples": [{"grid": [n_, n_], "seconds": t_} for n_, t_ in samples]},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def metric_name(args):
    return METRIC if args.n == 4096 else f"solves_per_sec_{args.n}x{args.n}_TM_to_1e-10"


def workload_config(args):
    return {"workload": f"synthetic TM device {args.n}x{args.n} (dh=0.02um=lambda0/75, Npml=15, eps=12 waveguide + seeded eps 2..12.25 "
                        f"cylinders/boxes at {args.density:.5f}/um^2, x-normal line source); one step = a sweep of {args.sweep} frequencies "
                        f"(200 THz + k*0.5 THz) per GPU solved concurrently, each to 1e-10 relative residual incl. per-frequency "
                        f"setup and H recovery; value counts solved (omega,source) right-hand sides per second",
            "grid": [args.n, args.n],
            "solver": ("BiCGSTAB + shifted-Laplacian multigrid (fp32) / fp64 operator" if args.solver == "bicgstab" or (args.solver == "auto" and args.n * args.n < 2 ** 22) else
                       f"multilevel Krylov (flexible GMRES per level, steps {args.ml_spec or '6,6'}, F cycles) + shifted-Laplacian multigrid (fp32) / fp64 operator"),
            "l2": "inputs larger than L2: one complex128 vector is 268 MB vs 126 MB L2",
            "parallelism": f"omega sweep: {args.sweep} frequencies per GPU per step, the same frequency set on every rank (independent replicas of the unit of work), no data-path collective"}


# ------------------------------------------------------------------------------------------------------------
def b200_arm(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import fdfd_jl_b200 as fdfd
    from importlib import import_module
    wl = import_module("fdfd_jl_b200.workloads")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()  # the library's main stream; worker streams of the sweep fork from the same device
    ctx = fdfd.Context(local, stream=stream.cuda_stream)
    n, B = args.n, args.sweep
    N = n * n

    d = wl.synthetic_tm_device(fdfd, n, n, density=args.density)
    g = d.grid
    gc = g.as_c()
    # B frequencies per step, solved concurrently by the library (one stream each); every rank gets the SAME set (weak scaling
    # with identical work per GPU -- the iteration count depends on the frequency)
    omegas = [2 * math.pi * (200e12 + 0.5e12 * k) for k in range(B)]
    wB = (C.c_double * B)(*omegas)
    opts = fdfd.default_opts(concurrency=args.concurrency or B)
    if args.solver != "auto":
        opts.solver = {"bicgstab": fdfd._lib.SOLVER_BICGSTAB, "mlkrylov": fdfd._lib.SOLVER_MLKRYLOV}[args.solver]
    if args.ml_spec or args.ml_restart:
        k = [int(x) for x in (args.ml_spec or "0").split(",")] + [0, 0, 0]
        opts.ml_spec = k[0] | (k[1] << 8) | (k[2] << 16) | (args.ml_restart << 24)
    # torch is plumbing: device memory and pinned host buffers
    eps_h = torch.from_numpy(np.asfortranarray(d.eps_r).ravel(order="F").copy()).pin_memory()
    src_h = torch.from_numpy(np.asfortranarray(d.src).ravel(order="F").copy()).pin_memory()
    eps_d = eps_h.cuda(non_blocking=True)
    src_d = src_h.cuda(non_blocking=True)
    fields_d = torch.empty(B * 3 * N, dtype=torch.complex128, device="cuda")
    fields_h = torch.empty(B * 3 * N, dtype=torch.complex128).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sweep(eps_ptr, src_ptr, out_ptr):
        """the hot path through the C ABI: fdfd_solve_driven == solve(d::Device, pol) for B frequencies"""
        infos = (fdfd.Info * B)()
        code = fdfd.lib().fdfd_solve_driven(ctx.handle, C.byref(gc), fdfd.TM, B, wB, fdfd.ptr(eps_ptr), fdfd.ptr(src_ptr), 0,
                                            C.byref(opts), fdfd.ptr(out_ptr), infos)
        fdfd.check(code, ctx.handle)
        return [i.asdict() for i in infos]

    # ---- device-resident timing: CUDA events on the library's stream, barrier + synchronize on both sides
    infos = []
    for _ in range(args.warmup):
        sweep(eps_d.data_ptr(), src_d.data_ptr(), fields_d.data_ptr())
    barrier()
    l0 = ctx.launch_count()
    with ClockSampler(local) as clk:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            infos += sweep(eps_d.data_ptr(), src_d.data_ptr(), fields_d.data_ptr())
            stream.synchronize()
        e1.record(stream)
        torch.cuda.synchronize()
        t_res = e0.elapsed_time(e1) * 1e-3
    launches = ctx.launch_count() - l0
    barrier()
    tt = torch.tensor([t_res], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt.item())

    # ---- end-to-end: the same call with pinned HOST buffers (H2D of eps_r/src, D2H of the B (Nx,Ny,3) fields inside)
    e2e_steps = max(1, min(args.steps, 2))
    sweep(eps_h.data_ptr(), src_h.data_ptr(), fields_h.data_ptr())
    barrier()
    e2e_infos = []
    with ClockSampler(local) as clk_e2e:   # the e2e leg runs after the resident one on a warmer GPU: its clocks are logged separately
        e2 = torch.cuda.Event(enable_timing=True); e3 = torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(e2e_steps):
            e2e_infos += sweep(eps_h.data_ptr(), src_h.data_ptr(), fields_h.data_ptr())
            stream.synchronize()
        e3.record(stream)
        torch.cuda.synchronize()
    te = torch.tensor([e2.elapsed_time(e3) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())

    # ---- roofline of the stencil apply kernel, timed live with CUDA events on the launching stream
    P = fdfd.Problem(g, fdfd.TM, omegas[0], eps_d.data_ptr(), ctx=ctx, precond=0)
    ms_apply = P.bench_apply(200)
    P.close()
    # the multigrid kernels that take most of a step's time, timed the same way on the level-0 arrays
    P = fdfd.Problem(g, fdfd.TM, omegas[0], eps_d.data_ptr(), ctx=ctx)
    mg_kernels = []
    for kind, name, bpp in ((0, "k_smooth3 + k_lines2 (fp32 smoothing sweep fused with the coarse-grid correction, level 0)", 34.0),
                            (1, "k_restrict_tile (fp32 residual + restriction, level 0)", 26.0),
                            (2, "k_smooth2<ZERO> + k_lines2 (fp32 zero-guess sweep, level 0)", 24.0)):
        ms_k = P.bench_mg(kind, 100)
        mg_kernels.append({"kernel": name, "algorithmic_bytes_per_point": bpp, "ms_per_launch": ms_k, "achieved_gbs": bpp * N / (ms_k * 1e-3) / 1e9})
    ms_cycle = P.bench_mg(4, 50)
    # the batched stencil: B right-hand sides sharing the operator in ONE launch (the solvers do not use it yet -- see DESIGN.md 7)
    batched = []
    for nb in (1, 4, 8):
        ms_b = P.bench_apply_batched(nb, 60)
        bytes_b = (32.0 * nb + 16.0) * N
        batched.append({"nrhs": nb, "ms_per_launch": ms_b, "algorithmic_bytes_per_point_per_rhs": (32.0 * nb + 16.0) / nb,
                        "achieved_gbs": bytes_b / (ms_b * 1e-3) / 1e9, "us_per_rhs": 1e3 * ms_b / nb})
    P.close()
    peak, peak_src = measured_peak()
    achieved = ALG_BYTES_PER_POINT * N / (ms_apply * 1e-3) / 1e9

    ok = all(i["flag"] == 0 and i["relres"] <= 1e-10 for i in infos)
    flags = torch.tensor([1 if ok else 0], device="cuda")
    per_rank = [{"rank": rank, "seconds": t_res, "iters": [i["iters"] for i in infos], "krylov_ms": [round(i["solve_ms"], 1) for i in infos]}]
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank[0])
        per_rank = gathered
    slab_rec = None
    sg = args.slab_grid if args.slab_grid >= 0 else (8192 if world > 1 else 0)
    if sg > 0:
        ctx_s = ctx
        try:
            slab_rec = slab_record(args, sg, fdfd, ctx_s, stream, rank, world, local)
        except Exception as e:  # noqa: BLE001 -- the sweep line must survive a failed slab leg
            slab_rec = {"error": str(e)[-300:]}
    if rank == 0:
        value = world * B * args.steps / t_max
        line = {"metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "c128", "data": "synthetic", "config": workload_config(args),
                "converged": bool(flags.item()),
                "solve": {"solves_per_step_per_gpu": B, "iters": [i["iters"] for i in infos], "relres_max": max(i["relres"] for i in infos),
                          "krylov_ms": [round(i["solve_ms"], 1) for i in infos], "setup_ms": [round(i["setup_ms"], 1) for i in infos],
                          "restarts": [i["restarts"] for i in infos], "mg_levels": infos[0]["mg_levels"], "per_rank": per_rank},
                "e2e": {"value": world * B * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": world * 2 * N * 16,
                        "d2h_bytes_per_step": world * B * 3 * N * 16, "steps": e2e_steps, "clocks": clk_e2e.summary(),
                        "call_ms": [round(i["total_ms"], 1) for i in e2e_infos], "krylov_ms": [round(i["solve_ms"], 1) for i in e2e_infos]},
                "gpu_launches": int(launches),
                "clocks": clk.summary(),
                "roofline": {"kernel": "k_apply (matrix-free complex128 Yee stencil, TM)", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": ALG_BYTES_PER_POINT * N, "ms_per_launch": ms_apply,
                             "stencil_hbm_gbs": achieved},
                "roofline_multigrid": {"note": "the kernels that take most of a step (k_apply is ~9 % of it); same live CUDA-event timing, same peak",
                                       "peak": peak, "unit": "GB/s",
                                       "kernels": [dict(k, frac=k["achieved_gbs"] / peak) for k in mg_kernels],
                                       "ms_per_cycle": ms_cycle},
                "roofline_batched": {"kernel": "k_apply_batched (B right-hand sides sharing one operator per launch: (32 B + 16) / B algorithmic B/pt/rhs; "
                                               "kernel-level measurement, the Krylov solvers take one right-hand side per solve)",
                                     "peak": peak, "unit": "GB/s", "runs": [dict(b, frac=b["achieved_gbs"] / peak) for b in batched],
                                     "speedup_per_rhs_B4_vs_single_k_apply": ms_apply / (batched[1]["ms_per_launch"] / 4.0)}}
        if slab_rec is not None:
            line["slab"] = slab_rec
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            t_cpu, _ = cpu_reference_solve(args.ref_n, args.density)
            t_half, _ = cpu_reference_solve(args.ref_n // 2, args.density)
            expo, how = fit_exponent([(args.ref_n // 2, t_half), (args.ref_n, t_cpu)])
            scale = (N / float(args.ref_n ** 2)) ** expo
            line["cpu_baseline"] = {"value": 1.0 / (t_cpu * scale), "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"oracle (SciPy SuperLU direct solve, stand-in for Julia \\/UMFPACK) on the same device, SOLVED at "
                                              f"{args.ref_n}^2: {t_cpu:.2f} s; extrapolated to {n}^2 by t ~ N^{expo:.3f} ({how})",
                                    "sample_seconds_per_solve": t_cpu, "sample_grid": [args.ref_n, args.ref_n], "fit_exponent": expo}
        # DRAM bytes per launch of the SAME k_apply instantiation (complex128 in, no fused dot, 4 rows per thread), from the
        # `ncu --set full` capture of this round (tools/r2_profile.sh; profiles/r02_k_apply_traffic.json says which command)
        prof = os.path.join(ROOT, "profiles", "r02_k_apply_traffic.json")
        if os.path.exists(prof) and n == 4096:
            try:
                line["roofline"]["traffic"] = json.load(open(prof))["dram_bytes_per_launch"]
                line["roofline"]["traffic_source"] = "profiles/r02_k_apply_traffic.json"
            except Exception:
                pass
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def slab_run(args, n, fdfd, ctx, stream, rank, world, local, steps, warmup, e2e):
    """ONE n x n TM solve split into row slabs over the ranks (SURVEY §8e second mode, BASELINE config 5):
    step = one fdfd_solve_driven_slab call on every rank.  Returns the record on every rank (rank 0's is printed)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from importlib import import_module
    wl = import_module("fdfd_jl_b200.workloads")
    slab = import_module("fdfd_jl_b200.slab")
    comm = slab.SlabComm.nccl(ctx, rank, world)
    g0 = fdfd.Grid(0.02, [15, 15], [0.0, n * 0.02], [0.0, n * 0.02])
    y0, nr = slab.slab_rows(g0, world, rank)
    g, omega, eps_rows, src_rows = wl.synthetic_tm_device(fdfd, n, n, density=args.density, rows=(y0, nr))
    gc = g.as_c()
    opts = fdfd.default_opts()
    if args.maxit > 0:
        opts.maxit = args.maxit
    M = n * nr
    eps_h = torch.from_numpy(np.asfortranarray(eps_rows).ravel(order="F").copy()).pin_memory()
    src_h = torch.from_numpy(np.asfortranarray(src_rows).ravel(order="F").copy()).pin_memory()
    eps_d, src_d = eps_h.cuda(), src_h.cuda()
    fields_d = torch.empty(3 * M, dtype=torch.complex128, device="cuda")
    fields_h = torch.empty(3 * M, dtype=torch.complex128).pin_memory() if e2e else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(e, s_, f):
        info = fdfd.Info()
        code = fdfd.lib().fdfd_solve_driven_slab(ctx.handle, comm.handle, C.byref(gc), omega, fdfd.ptr(e), fdfd.ptr(s_), C.byref(opts),
                                                 fdfd.ptr(f), C.byref(info))
        if code not in (0, fdfd._lib.ERR_NOCONV, fdfd._lib.ERR_BREAKDOWN):   # an unconverged solve is reported, not raised: every rank sees the same flag
            fdfd.check(code, ctx.handle)
        return info.asdict()

    infos = []
    for _ in range(warmup):
        step(eps_d.data_ptr(), src_d.data_ptr(), fields_d.data_ptr())
    barrier()
    l0 = ctx.launch_count()
    with ClockSampler(local) as clk:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            infos.append(step(eps_d.data_ptr(), src_d.data_ptr(), fields_d.data_ptr()))
            stream.synchronize()
        e1.record(stream)
        torch.cuda.synchronize()
        t_res = e0.elapsed_time(e1) * 1e-3
    launches = ctx.launch_count() - l0
    barrier()
    e2 = torch.cuda.Event(enable_timing=True); e3 = torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    if e2e:
        step(eps_h.data_ptr(), src_h.data_ptr(), fields_h.data_ptr())
    stream.synchronize()
    e3.record(stream)
    torch.cuda.synchronize()
    tt = torch.tensor([t_res, e2.elapsed_time(e3) * 1e-3], dtype=torch.float64, device="cuda")
    ok = torch.tensor([1 if all(i["flag"] == 0 and i["relres"] <= 1e-10 for i in infos) else 0], device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    t_max, t_e2e = float(tt[0].item()), float(tt[1].item())
    st = comm.stats()
    i0 = infos[-1]
    rec = {"metric": f"solves_per_sec_{n}x{n}_TM_slab_to_1e-10", "value": steps / t_max, "unit": UNIT, "n_gpus": world,
           "steps": steps, "warmup": warmup, "ms_per_step": t_max / steps * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
           "config": {"workload": f"ONE synthetic TM device {n}x{n} (same map family as the sweep workload) split into {world} row slabs of "
                                  f"{nr} rows; step = setup + BiCGSTAB/multigrid to 1e-10 + H recovery on every slab",
                      "grid": [n, n], "parallelism": f"y-slabs x{world}: ring halo exchange after every stencil-type kernel, "
                                                     "one 4-double allreduce per dot product, iteration replayed as one CUDA graph",
                      "l2": "inputs larger than L2"},
           "converged": bool(ok.item()),
           "solve": {"iters": [i["iters"] for i in infos], "relres": [i["relres"] for i in infos],
                     "krylov_ms": [round(i["solve_ms"], 1) for i in infos], "setup_ms": [round(i["setup_ms"], 1) for i in infos],
                     "restarts": [i["restarts"] for i in infos],
                     "mg_levels": i0["mg_levels"], "ms_per_iteration": i0["solve_ms"] / max(1, i0["iters"])},
           "e2e": None if not e2e else {"value": 1.0 / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 2 * M * 16 * world,
                                        "d2h_bytes_per_step": 3 * M * 16 * world, "steps": 1},
           "gpu_launches": int(launches), "comm_per_rank_captured": st, "clocks": clk.summary()}
    comm.close()
    return rec


def slab_record(args, n, fdfd, ctx, stream, rank, world, local):
    """the slab strong-scaling datum carried by the sweep line: one solve, no warm-up repetition (a solve is tens of seconds)"""
    r = slab_run(args, n, fdfd, ctx, stream, rank, world, local, steps=1, warmup=0, e2e=False)
    return {k: r[k] for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "scaling", "converged", "solve", "config", "comm_per_rank_captured", "gpu_launches")}


def slab_arm(args):
    import torch
    import torch.distributed as dist
    import fdfd_jl_b200 as fdfd
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    ctx = fdfd.Context(local, stream=stream.cuda_stream)
    rec = slab_run(args, args.n, fdfd, ctx, stream, rank, world, local, steps=args.steps, warmup=args.warmup, e2e=not args.no_e2e)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    elif a.mode == "slab":
        slab_arm(a)
    else:
        b200_arm(a)
