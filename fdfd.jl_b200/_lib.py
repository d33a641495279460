"""ctypes binding of libfdfd_b200.so (include/fdfd_b200.h).  Same symbols a Julia `ccall` wrapper binds
(see INTEGRATION.md).  There is no CPU fallback: a missing library or a missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfdfd_b200.so")

TM, TE = 1, 2
ORDER_FB, ORDER_BF = 0, 1
DXF, DXB, DYF, DYB = 0, 1, 2, 3
CSR, CSC = 0, 1
SOLVER_BICGSTAB, SOLVER_COCG, SOLVER_MLKRYLOV, SOLVER_AUTO = 0, 1, 2, 3
PRECOND_NONE, PRECOND_JACOBI, PRECOND_MG = 0, 1, 2
MG_F32, MG_F64 = 0, 1
CYCLE_V, CYCLE_F, CYCLE_W = 0, 1, 2
WHICH = {"LM": 0, "LR": 1, "SR": 2, "LI": 3, "SI": 4}
OK, ERR_ARG, ERR_CUDA, ERR_NOCONV, ERR_BREAKDOWN, ERR_ALLOC = range(6)


class GridT(C.Structure):
    _fields_ = [("Nx", C.c_int64), ("Ny", C.c_int64), ("Npml_x", C.c_int64), ("Npml_y", C.c_int64),
                ("x0", C.c_double), ("x1", C.c_double), ("y0", C.c_double), ("y1", C.c_double), ("L0", C.c_double)]


class SolveOpts(C.Structure):
    _fields_ = [("solver", C.c_int32), ("precond", C.c_int32), ("tol", C.c_double), ("maxit", C.c_int32),
                ("mg_precision", C.c_int32), ("mg_cycle", C.c_int32), ("mg_wdepth", C.c_int32),
                ("mg_nu1", C.c_int32), ("mg_nu2", C.c_int32), ("mg_coarse_sweeps", C.c_int32),
                ("mg_beta", C.c_double), ("mg_wjac", C.c_double), ("mg_wline", C.c_double),
                ("check_every", C.c_int32), ("verbose", C.c_int32),
                ("mg_shift_growth", C.c_double), ("mg_max_levels", C.c_int32), ("use_graph", C.c_int32),
                ("concurrency", C.c_int32), ("ml_spec", C.c_int32)]


class Info(C.Structure):
    _fields_ = [("iters", C.c_int32), ("flag", C.c_int32), ("relres", C.c_double), ("setup_ms", C.c_double),
                ("solve_ms", C.c_double), ("total_ms", C.c_double), ("launches", C.c_int64),
                ("restarts", C.c_int32), ("mg_levels", C.c_int32)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class FdfdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fdfd_b200 error {code}: {msg}")
        self.code = code


_lib = None

# every exported symbol of include/fdfd_b200.h (tests check the library exports all of them)
EXPORTS = [
    "fdfd_abi_version", "fdfd_ctx_create", "fdfd_ctx_destroy", "fdfd_last_error", "fdfd_launch_count",
    "fdfd_default_opts", "fdfd_sfactors", "fdfd_assemble_derivative", "fdfd_assemble_system",
    "fdfd_apply_operator", "fdfd_solve_driven", "fdfd_solve_modulated", "fdfd_eigenfrequency",
    "fdfd_problem_create", "fdfd_problem_destroy", "fdfd_problem_set_rhs", "fdfd_problem_set_source",
    "fdfd_problem_solve", "fdfd_problem_get_solution", "fdfd_problem_get_fields", "fdfd_problem_bench_apply", "fdfd_problem_bench_mg",
    "fdfd_problem_precond", "fdfd_problem_get_history", "fdfd_debug_hess_eig",
    "fdfd_debug_general_eig", "fdfd_debug_krylov_schur", "fdfd_apply_operator_batched", "fdfd_problem_bench_apply_batched",
    "fdfd_problem_flux_x", "fdfd_rasterize",
    "fdfd_comm_unique_id", "fdfd_comm_create_nccl", "fdfd_comm_group_create", "fdfd_comm_group_destroy",
    "fdfd_comm_create_threads", "fdfd_comm_destroy", "fdfd_slab_rows", "fdfd_solve_driven_slab", "fdfd_comm_stats",
    "fdfd_solve_modulated_slab", "fdfd_eigenfrequency_slab", "fdfd_problem_ml_cycles",
    "fdfd_debug_ml_lsq", "fdfd_debug_ml_transfer", "fdfd_debug_ml_lsq_gpu", "fdfd_debug_ml_transfer_gpu",
    "fdfd_dolinearsolve_csc", "fdfd_dolinearsolve_csc_grid", "fdfd_debug_sell_spmv", "fdfd_debug_sell_bench",
]
COMM_THREADS, COMM_NCCL = 0, 1
COMM_ID_BYTES = 128


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FdfdError(-1, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(make -C fdfd.jl_b200/csrc).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.fdfd_abi_version.restype = i32
        L.fdfd_last_error.restype = C.c_char_p
        L.fdfd_last_error.argtypes = [vp]
        L.fdfd_launch_count.restype = i64
        L.fdfd_launch_count.argtypes = [vp]
        L.fdfd_ctx_create.argtypes = [i32, vp, C.POINTER(vp)]
        L.fdfd_ctx_destroy.argtypes = [vp]
        L.fdfd_ctx_destroy.restype = None
        L.fdfd_default_opts.argtypes = [C.POINTER(SolveOpts)]
        L.fdfd_default_opts.restype = None
        G = C.POINTER(GridT)
        L.fdfd_sfactors.argtypes = [vp, G, dbl, vp, vp, vp, vp]
        L.fdfd_assemble_derivative.argtypes = [vp, G, dbl, i32, i32, i32, i32, vp, vp, vp]
        L.fdfd_assemble_system.argtypes = [vp, G, i32, i32, dbl, vp, i32, i32, vp, vp, vp]
        L.fdfd_apply_operator.argtypes = [vp, G, i32, i32, dbl, vp, vp, vp]
        L.fdfd_solve_driven.argtypes = [vp, G, i32, i32, C.POINTER(dbl), vp, vp, i32, C.POINTER(SolveOpts), vp, C.POINTER(Info)]
        L.fdfd_solve_modulated.argtypes = [vp, G, dbl, dbl, i32, i32, vp, vp, vp, C.POINTER(SolveOpts), vp, C.POINTER(Info)]
        L.fdfd_eigenfrequency.argtypes = [vp, G, i32, dbl, i32, i32, i32, vp, C.POINTER(SolveOpts), vp, vp, C.POINTER(Info)]
        L.fdfd_problem_create.argtypes = [vp, G, i32, i32, dbl, vp, C.POINTER(SolveOpts), C.POINTER(vp)]
        L.fdfd_problem_destroy.argtypes = [vp]
        L.fdfd_problem_destroy.restype = None
        L.fdfd_problem_set_rhs.argtypes = [vp, vp]
        L.fdfd_problem_set_source.argtypes = [vp, vp]
        L.fdfd_problem_solve.argtypes = [vp, C.POINTER(Info)]
        L.fdfd_problem_get_solution.argtypes = [vp, vp]
        L.fdfd_problem_get_fields.argtypes = [vp, i32, vp]
        L.fdfd_problem_bench_apply.argtypes = [vp, i32, C.POINTER(dbl)]
        L.fdfd_problem_bench_mg.argtypes = [vp, i32, i32, C.POINTER(dbl)]
        L.fdfd_problem_precond.argtypes = [vp, vp, vp]
        L.fdfd_problem_ml_cycles.argtypes = [vp, C.POINTER(i64)]
        L.fdfd_debug_ml_lsq.argtypes = [i32, vp, dbl, vp, C.POINTER(dbl)]
        L.fdfd_debug_ml_transfer.argtypes = [i64, i64, i32, dbl, vp, vp]
        L.fdfd_debug_ml_lsq_gpu.argtypes = [vp, i32, vp, dbl, vp, C.POINTER(dbl)]
        L.fdfd_debug_ml_transfer_gpu.argtypes = [vp, i64, i64, i32, dbl, vp, vp]
        L.fdfd_problem_get_history.argtypes = [vp, vp, i32, C.POINTER(i32)]
        L.fdfd_debug_hess_eig.argtypes = [i32, vp, vp, vp]
        L.fdfd_apply_operator_batched.argtypes = [vp, G, i32, i32, dbl, vp, i32, vp, vp]
        L.fdfd_problem_bench_apply_batched.argtypes = [vp, i32, i32, C.POINTER(dbl)]
        L.fdfd_debug_general_eig.argtypes = [i32, vp, vp, vp]
        L.fdfd_debug_krylov_schur.argtypes = [i32, vp, i32, i32, i32, dbl, i32, vp, vp, C.POINTER(i32), C.POINTER(i32)]
        L.fdfd_problem_flux_x.argtypes = [vp, dbl, dbl, dbl, i32, C.POINTER(dbl)]
        L.fdfd_rasterize.argtypes = [vp, G, i32, vp, vp]
        L.fdfd_comm_unique_id.argtypes = [vp]
        L.fdfd_comm_create_nccl.argtypes = [vp, i32, i32, vp, C.POINTER(vp)]
        L.fdfd_comm_group_create.argtypes = [i32, C.POINTER(vp)]
        L.fdfd_comm_group_destroy.argtypes = [vp]
        L.fdfd_comm_group_destroy.restype = None
        L.fdfd_comm_create_threads.argtypes = [vp, i32, C.POINTER(vp)]
        L.fdfd_comm_destroy.argtypes = [vp]
        L.fdfd_comm_destroy.restype = None
        L.fdfd_slab_rows.argtypes = [G, i32, i32, C.POINTER(i64), C.POINTER(i64)]
        L.fdfd_solve_driven_slab.argtypes = [vp, vp, G, dbl, vp, vp, C.POINTER(SolveOpts), vp, C.POINTER(Info)]
        L.fdfd_solve_modulated_slab.argtypes = [vp, vp, G, dbl, dbl, i32, i32, vp, vp, vp, C.POINTER(SolveOpts), vp, C.POINTER(Info)]
        L.fdfd_eigenfrequency_slab.argtypes = [vp, vp, G, i32, dbl, i32, i32, i32, vp, C.POINTER(SolveOpts), vp, vp, C.POINTER(Info)]
        L.fdfd_dolinearsolve_csc.argtypes = [vp, i64, vp, vp, vp, i32, vp, C.POINTER(SolveOpts), vp, C.POINTER(Info)]
        L.fdfd_dolinearsolve_csc_grid.argtypes = [vp, G, dbl, i64, vp, vp, vp, i32, vp, C.POINTER(SolveOpts), vp, C.POINTER(Info)]
        L.fdfd_debug_sell_bench.argtypes = [vp, i64, vp, vp, vp, i32, i32, C.POINTER(dbl), C.POINTER(dbl)]
        L.fdfd_debug_sell_spmv.argtypes = [i64, vp, vp, vp, i32, vp, vp, vp, vp, C.POINTER(i64)]
        L.fdfd_comm_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
        _lib = L
    return _lib


def check(code, ctx=None):
    if code != OK:
        msg = lib().fdfd_last_error(ctx).decode("utf-8", "replace")
        raise FdfdError(code, msg)


def default_opts(**kw) -> SolveOpts:
    o = SolveOpts()
    lib().fdfd_default_opts(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown solver option {k!r}")
        setattr(o, k, v)
    return o


def as_c128(a, shape=None):
    """Column-major (x fastest) complex128 buffer for an (Nx,Ny[,..]) array; Julia's `ComplexF64.(a)`."""
    a = np.asarray(a, dtype=np.complex128)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return np.asfortranarray(a)


def ptr(a):
    """void* of a numpy array, or pass through ints (device pointers, e.g. torch tensor.data_ptr())."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One per GPU (one per process in multi-GPU runs)."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        code = lib().fdfd_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h))
        check(code, None)
        self.device = device

    @property
    def handle(self):
        return self._h

    def launch_count(self):
        return int(lib().fdfd_launch_count(self._h))

    def close(self):
        if self._h:
            lib().fdfd_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        dev = int(os.environ.get("LOCAL_RANK", "0")) if "FDFD_B200_DEVICE" not in os.environ else int(os.environ["FDFD_B200_DEVICE"])
        _default_ctx = Context(dev)
    return _default_ctx
