"""TEST / BENCH INFRASTRUCTURE -- ctypes binding of oracle/_build/libriptrm_det.so (oracle/c/riptrm_det.c).
Nothing on the product path imports this."""
import ctypes as C
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import build as _build

TRACE_FIELDS, SUMMARY_FIELDS = 25, 16
_DP = C.POINTER(C.c_double)
_lib = None


class DetOptions(C.Structure):
    _fields_ = [
        ("maxiter", C.c_int32), ("inner_maxiter", C.c_int32), ("tcg_mininner", C.c_int32),
        ("tcg_maxinner", C.c_int32), ("is_euclidean_embedded", C.c_int32), ("trace_mode", C.c_int32),
        ("trace_capacity", C.c_int32), ("schedule_split", C.c_int32),
        ("tolresid", C.c_double), ("maxtime", C.c_double), ("inner_maxtime", C.c_double),
        ("initial_tr_radius", C.c_double), ("minimal_initial_tr_radius", C.c_double),
        ("maximal_tr_radius", C.c_double), ("rho", C.c_double), ("reduction_regularization", C.c_double),
        ("gamma", C.c_double), ("const_left", C.c_double), ("const_right", C.c_double),
        ("tcg_theta", C.c_double), ("tcg_kappa", C.c_double),
        ("mu_sched", _DP), ("tol_lagrangian_sched", _DP), ("tol_complementarity_sched", _DP),
    ]


def available():
    return os.path.exists(_build.LIB) or _try_build()


def _try_build():
    try:
        _build.build()
        return True
    except Exception:
        return False


_lib_faithful = None


def load(faithful=False):
    """faithful=False: the merged-reduction tCG (bit-identical to the CUDA kernel); True: reference operation order."""
    global _lib, _lib_faithful
    if _lib is None:
        _build.build()
        _lib = C.CDLL(_build.LIB)
        _lib_faithful = C.CDLL(_build.LIB_FAITHFUL)
        for l in (_lib, _lib_faithful):
            l.riptrm_det_solve_nonnegpca.restype = C.c_int
            l.riptrm_det_solve_many.restype = C.c_int
            l.riptrm_det_hessvec.restype = C.c_int
    return _lib_faithful if faithful else _lib


def make_options(option=None, trace_mode=0, trace_capacity=0):
    """RIPTRM.py:305-358 defaults (oracle/riptrm_oracle.default_option) merged under `option`; the callable
    keys are evaluated into per-outer-iteration schedules (RIPTRM.py:881-885, :890-893)."""
    from oracle.riptrm_oracle import default_option
    o = default_option()
    o.update(option or {})
    K = int(o["maxiter"])
    mu = [float(o["initial_barrier_parameter"])]
    for _ in range(K):
        m = mu[-1]
        r_, c_, b_ = o["barrier_parameter_update_r"], o["barrier_parameter_update_c"], o["barrier_parameter_update_b"]
        if o["do_simple_barrier_parameter_update"]:
            mu.append(max(o["min_barrier_parameter"], c_ * (m ** (1 + r_))))
        else:
            mu.append(max(o["min_barrier_parameter"], min(b_ * m, c_ * (m ** (1 + r_)))))
    mu = np.array(mu)
    tolL = np.array([o["forcing_function_Lagrangian"](float(m)) for m in mu])
    tolC = np.array([o["forcing_function_complementarity"](float(m)) for m in mu])
    d = DetOptions()
    d.maxiter = K
    d.inner_maxiter = -1 if o["inner_maxiter"] is None else int(o["inner_maxiter"])
    d.tcg_mininner = int(o["tCG_mininner"])
    d.tcg_maxinner = -1
    d.is_euclidean_embedded = int(bool(o["is_euclidean_embedded"]))
    d.trace_mode, d.trace_capacity = int(trace_mode), int(trace_capacity)
    d.tolresid = float(o["tolresid"])
    d.maxtime, d.inner_maxtime = 1e300, -1.0
    d.initial_tr_radius = -1.0 if o["initial_TR_radius"] is None else float(o["initial_TR_radius"])
    d.minimal_initial_tr_radius = float(o["minimal_initial_TR_radius"])
    d.maximal_tr_radius = float(o["maximal_TR_radius"])
    d.rho, d.reduction_regularization, d.gamma = float(o["rho"]), float(o["reduction_regularization"]), float(o["gamma"])
    d.const_left, d.const_right = float(o["const_left"]), float(o["const_right"])
    d.tcg_theta, d.tcg_kappa = float(o["tCG_theta"]), float(o["tCG_kappa"])
    d.mu_sched = mu.ctypes.data_as(_DP)
    d.tol_lagrangian_sched = tolL.ctypes.data_as(_DP)
    d.tol_complementarity_sched = tolC.ctypes.data_as(_DP)
    return d, (mu, tolL, tolC)


def _p(a):
    return a.ctypes.data_as(_DP) if a is not None else None


def solve(Z, x0, y0, option=None, eps=0.0, trace_capacity=0, faithful=False):
    """One NonnegPCA/Sphere solve.  Returns (x, y, summary[16], trace[rows, 25] | None)."""
    lib = load(faithful)
    Z, x0, y0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (Z, x0, y0))
    n = x0.shape[0]
    d, keep = make_options(option, 1 if trace_capacity else 0, trace_capacity)
    x, y, sm = np.empty(n), np.empty(n), np.empty(SUMMARY_FIELDS)
    tr = np.full((trace_capacity, TRACE_FIELDS), np.nan) if trace_capacity else None
    rc = lib.riptrm_det_solve_nonnegpca(C.c_int(n), _p(Z), _p(x0), _p(y0), C.c_double(eps), C.byref(d), _p(x), _p(y),
                                        _p(sm), _p(tr))
    if rc:
        raise RuntimeError(f"riptrm_det_solve_nonnegpca rc={rc}")
    if tr is not None:
        rows = int(sm[15])
        if rows > trace_capacity:
            return solve(Z, x0, y0, option, eps, rows, faithful)
        tr = tr[:rows]
    return x, y, sm, tr


def hessvec(Z, x, y, mu, v, eps=0.0):
    lib = load()
    Z, x, y, v = (np.ascontiguousarray(a, dtype=np.float64) for a in (Z, x, y, v))
    out = np.empty_like(x)
    rc = lib.riptrm_det_hessvec(C.c_int(x.shape[0]), _p(Z), _p(x), _p(y), C.c_double(eps), C.c_double(mu), _p(v), _p(out))
    if rc:
        raise RuntimeError(f"riptrm_det_hessvec rc={rc}")
    return out


def solve_many(Z, x0, y0, option=None, eps=0.0, threads=1):
    """Z [B, n, n], x0/y0 [B, n] -> (x, y, summary [B, 16]); `threads` host threads (ctypes drops the GIL)."""
    lib = load()
    Z, x0, y0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (Z, x0, y0))
    B, n = x0.shape
    d, keep = make_options(option)
    x, y, sm = np.empty((B, n)), np.empty((B, n)), np.empty((B, SUMMARY_FIELDS))
    threads = max(1, min(threads, B))
    bounds = np.linspace(0, B, threads * 4 + 1).astype(int) if threads > 1 else np.array([0, B])

    def work(i):
        lo, hi = int(bounds[i]), int(bounds[i + 1])
        if hi <= lo:
            return 0
        return lib.riptrm_det_solve_many(C.c_int(hi - lo), C.c_int(n), _p(Z[lo:hi]), _p(x0[lo:hi]), _p(y0[lo:hi]),
                                         C.c_double(eps), C.byref(d), _p(x[lo:hi]), _p(y[lo:hi]), _p(sm[lo:hi]))
    if threads == 1:
        rcs = [work(0)]
    else:
        with ThreadPoolExecutor(threads) as ex:
            rcs = list(ex.map(work, range(len(bounds) - 1)))
    if any(rcs):
        raise RuntimeError("riptrm_det_solve_many failed")
    return x, y, sm


def run_sample(protocol, dim, first_seed, target_seconds, threads, max_pairs, points_per_instance=1):
    """bench.py's CPU leg: the first pairs of the bench sweep (instances first_seed.., `points_per_instance` initial
    points each) on `threads` host threads, sized for about `target_seconds`."""
    from oracle.problems import nonnegpca_generate_sweep
    opt = {k: v for k, v in protocol.items() if k not in ("TRS_solver", "second_order_stationarity", "maxtime")}
    ipp = points_per_instance

    def gen(n_inst):
        Z, X, Y = nonnegpca_generate_sweep(first_seed, n_inst, ipp, dim)
        return np.repeat(Z, ipp, axis=0), X, Y
    ncal = max(1, min(max_pairs, 2 * threads) // ipp)
    Zc, Xc, Yc = gen(ncal)
    t = time.perf_counter()
    solve_many(Zc, Xc, Yc, opt, threads=threads)
    rate = ncal * ipp / (time.perf_counter() - t)
    n_inst = int(min(max_pairs // ipp, max(ncal, rate * target_seconds / ipp)))
    Z, X, Y = gen(n_inst)
    t = time.perf_counter()
    x, y, sm = solve_many(Z, X, Y, opt, threads=threads)
    secs = time.perf_counter() - t
    n = n_inst * ipp
    return {"pairs": n, "seconds": secs, "tcg_iters": int(sm[:, 12].sum()), "threads": threads, "kind": "port",
            "engine": "c",
            "sample": f"{n} pairs (instances {first_seed}..{first_seed + n_inst - 1} x {ipp} initial points) of the bench "
                      f"workload, C oracle (oracle/c/riptrm_det.c, closed-form derivatives, gcc -O2), {threads} host "
                      "threads, full protocol",
            "max_residual": float(sm[:, 1].max())}
