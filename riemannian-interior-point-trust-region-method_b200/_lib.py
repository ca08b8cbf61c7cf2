"""ctypes binding of include/riptrm_b200.h.  Fails loudly when the CUDA library is missing:
there is no CPU fallback on the product path."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# RIPTRM_B200_LIB: an alternative build of the same library (A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("RIPTRM_B200_LIB") or os.path.join(HERE, "csrc", "libriptrm_b200.so")

HOST, DEVICE = 0, 1
FAMILY_NONNEGPCA_SPHERE = 1
FAMILY_ROSENBROCK_GRASSMANN = 2
FAMILY_STABLEID_PRODUCT = 3
FAMILY_NONNEGPCA_COLUMNS = 4
FAMILY_NONNEGPCA_STIEFEL = 5
TRACE_FIELDS = 26
SUMMARY_FIELDS = 16

TCG_STOP_NAMES = ("MAX_INNER_ITER", "NEGATIVE_CURVATURE", "EXCEEDED_TR", "MODEL_INCREASED",
                  "REACHED_TARGET_LINEAR", "REACHED_TARGET_SUPERLINEAR")
# dxtype of the exact trust-region solver (TRSgep's `type`, RIPTRM.py:262-298): trace codes 6..11
TRS_TYPE_NAMES = {6: "boundary", 7: "interior", 8: "hardcase_1", 9: "hardcase_3", 10: "hardcase_6", 11: "hardcase_9"}
TRS_SOLVER_TCG, TRS_SOLVER_EXACT_REPMAT = 0, 1
INNER_STATUS_NAMES = (None, "converged", "primal_infeasible", "successful", "unsuccessful",
                      "max-time-exceeded", "max-iter-exceeded")
RADIUS_UPDATE_NAMES = (None, "reduced", "expanded", "unchanged")
STOP_REASONS = ("running", "maxtime", "maxiter", "tolresid", "numerical")

TR = {name: i for i, name in enumerate((
    "iteration", "num_inner", "mu", "TR_radius", "dxtype", "tcg_iters", "normdx", "minxfeasi", "minyfeasi",
    "compl", "ared/pred", "radius_update", "inner_status", "dual_clipping", "maxabsLagmult", "cost",
    "distance", "residual", "gradnorm", "complviolation", "dualviolation", "manviolation", "maxviolation",
    "meanviolation", "time", "mineigvalHw"))}
SM = {name: i for i, name in enumerate((
    "cost", "residual", "gradnorm", "complviolation", "dualviolation", "manviolation", "maxviolation",
    "meanviolation", "mu", "TR_radius", "outer_iters", "inner_iters", "tcg_iters", "aux_hessvecs",
    "stop_reason", "trace_rows"))}


class RiptrmOptions(C.Structure):
    _fields_ = [
        ("maxiter", C.c_int32), ("inner_maxiter", C.c_int32), ("tcg_mininner", C.c_int32),
        ("tcg_maxinner", C.c_int32), ("is_euclidean_embedded", C.c_int32), ("trace_mode", C.c_int32),
        ("trace_capacity", C.c_int32), ("schedule_split", C.c_int32),
        ("tolresid", C.c_double), ("maxtime", C.c_double), ("inner_maxtime", C.c_double),
        ("initial_tr_radius", C.c_double), ("minimal_initial_tr_radius", C.c_double),
        ("maximal_tr_radius", C.c_double), ("rho", C.c_double), ("reduction_regularization", C.c_double),
        ("gamma", C.c_double), ("const_left", C.c_double), ("const_right", C.c_double),
        ("tcg_theta", C.c_double), ("tcg_kappa", C.c_double),
        ("mu_sched", C.POINTER(C.c_double)), ("tol_lagrangian_sched", C.POINTER(C.c_double)),
        ("tol_complementarity_sched", C.POINTER(C.c_double)),
        ("trs_solver", C.c_int32), ("second_order_stationarity", C.c_int32), ("trs_tolhardcase", C.c_double),
        ("tol_second_order_sched", C.POINTER(C.c_double)),
    ]


class RiptrmError(RuntimeError):
    pass


_DP = C.POINTER(C.c_double)
_lib = None

# every symbol include/riptrm_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "riptrm_abi_version": (C.c_int, []),
    "riptrm_last_error": (C.c_char_p, []),
    "riptrm_create": (C.c_int, [C.c_int] * 6 + [C.POINTER(C.c_void_p)]),
    "riptrm_destroy": (C.c_int, [C.c_void_p]),
    "riptrm_set_nonnegpca": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "riptrm_set_rosenbrock": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "riptrm_set_stableid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p,
                                      C.c_int, C.c_int]),
    "riptrm_set_options": (C.c_int, [C.c_void_p, C.POINTER(RiptrmOptions)]),
    "riptrm_solve": (C.c_int, [C.c_void_p] * 7 + [C.c_int, C.c_void_p]),
    "riptrm_hessvec": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_void_p]),
    "riptrm_tcg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                             C.c_void_p, C.c_int, C.c_void_p]),
    "riptrm_trs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                             C.c_void_p, C.c_int, C.c_void_p]),
    "riptrm_newton": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "riptrm_trs_dense": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_void_p]),
    "riptrm_generate_nonnegpca": (C.c_int, [C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_double, C.c_double,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "riptrm_lane_placement": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "riptrm_launch_count": (C.c_int64, [C.c_void_p]),
    "riptrm_matvec_passes": (C.c_int64, [C.c_void_p]),
    "riptrm_last_kernel_ms": (C.c_double, [C.c_void_p]),
    "riptrm_measure_fp64_peaks": (C.c_int, [C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]),
}


def load_library():
    """dlopen()s the in-tree CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RiptrmError(
            f"{LIB_PATH} is missing: build it with `python {os.path.join(HERE, 'build.py')}` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.riptrm_abi_version() != 2:
        raise RiptrmError("libriptrm_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def measure_fp64_peaks(device=0, ms_target=20.0, repeats=3, stream=None):
    """(DFMA TFLOP/s, DMMA TFLOP/s) measured on `device` by the library's saturating micro-kernels."""
    out = (C.c_double * 2)()
    check(load_library().riptrm_measure_fp64_peaks(int(device), float(ms_target), int(repeats), out,
                                                   C.c_void_p(stream) if stream else None))
    return float(out[0]), float(out[1])


def trs_dense(A, a, Delta, tolhardcase=1e-8, device=0):
    """Batched dense trust-region subproblems on the GPU (`TRSgep` with B = I, RIPTRM.py:218-299): A [count, d, d] symmetric,
    a [count, d] -> (x [count, d], info [count, 4] = {type code, lam1, ||x||, smallest eigenvalue})."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    count, d = a.shape
    x, info = np.empty((count, d)), np.empty((count, 4))
    check(load_library().riptrm_trs_dense(int(device), d, count, ptr(A), ptr(a), float(Delta), float(tolhardcase), ptr(x),
                                          ptr(info), HOST, None))
    return x, info


def check(rc):
    if rc != 0:
        msg = load_library().riptrm_last_error()
        raise RiptrmError(f"riptrm error {rc}: {msg.decode() if msg else ''}")


def ptr(a):
    """void* of a C-contiguous float64 ndarray, a torch tensor, an int address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))
