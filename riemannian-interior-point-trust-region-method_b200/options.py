"""RIPTRM option handling on the host: the reference's defaults (src/solver/RIPTRM.py:305-358),
merged under the caller's dict exactly as RIPTRM.__init__ does (:359-361), and the evaluation of
the callable keys into per-outer-iteration schedules for the device (include/riptrm_b200.h)."""
import ctypes as C
import math

import numpy as np

from . import _lib


def default_option():
    """Same keys and values as RIPTRM.py:305-358.  `basisfun`: the reference's default draws a RANDOM tangent basis
    (utils.tangentorthobasis); the device builds a fixed orthonormal basis per manifold (basis.deterministic_basisfun is
    the same construction in NumPy) -- step and eigenvalues do not depend on the basis, so the key is accepted and the
    callable is not evaluated."""
    from .basis import deterministic_basisfun
    return {
        # Stopping criteria
        "maxtime": 240,
        "maxiter": 100,
        "tolresid": 1e-15,
        "inner_maxiter": None,
        "inner_maxtime": None,
        # Inner iteration setting
        "initial_TR_radius": None,
        "minimal_initial_TR_radius": 1e-15,
        "maximal_TR_radius": 10,
        "rho": 0.1,
        "reduction_regularization": 1e3,
        "gamma": 0.25,
        "forcing_function_Lagrangian": lambda mu: max(mu, 1e-14),
        "forcing_function_complementarity": lambda mu: max(1e-3 * mu, 1e-14),
        "forcing_function_second_order": lambda mu: mu,
        "min_barrier_parameter": 1e-15,
        "TRS_solver": "Exact_RepMat",
        "second_order_stationarity": True,
        "do_euclidean_lincomb": False,
        "is_euclidean_embedded": False,
        "TRS_tolresid": 1e-12,
        "TRS_tolhardcase": 1e-8,
        "tCG_theta": 1,
        "tCG_kappa": 0.1,
        "tCG_mininner": 1,
        "checkTRSoptimality": False,
        "initial_barrier_parameter": 0.1,
        "barrier_parameter_update_r": 0.01,
        "barrier_parameter_update_c": 0.5,
        "barrier_parameter_update_b": 0.8,
        "do_simple_barrier_parameter_update": True,
        "const_left": 0.5,
        "const_right": 1e20,
        "basisfun": deterministic_basisfun,
        # Display setting
        "verbosity": 0,
        "manviofun": lambda problem, x: 0,
        "callbackfun": lambda problem, x, y, z, eval: eval,
        # logging
        "save_inner_iteration": True,
        "wandb_logging": False,
        # Exit on error
        "do_exit_on_error": True,
    }


def barrier_schedule(option):
    """mu_sched[k] = barrier parameter of outer iteration k+1 (RIPTRM.py:852, :890-893), and the
    forcing-function values for it (:881-885).  Length maxiter + 1."""
    maxiter = int(option["maxiter"])
    r = option["barrier_parameter_update_r"]
    c = option["barrier_parameter_update_c"]
    b = option["barrier_parameter_update_b"]
    mu_min = option["min_barrier_parameter"]
    simple = option["do_simple_barrier_parameter_update"]
    fL = option["forcing_function_Lagrangian"]
    fC = option["forcing_function_complementarity"]
    mu = [float(option["initial_barrier_parameter"])]
    for _ in range(maxiter):
        m = mu[-1]
        if simple:
            mu.append(max(mu_min, c * (m ** (1 + r))))
        else:
            mu.append(max(mu_min, min(b * m, c * (m ** (1 + r)))))
    mu = np.array(mu, dtype=np.float64)
    tolL = np.array([fL(float(m)) for m in mu], dtype=np.float64)
    tolC = np.array([fC(float(m)) for m in mu], dtype=np.float64)
    return mu, tolL, tolC


def second_order_schedule(option, mu):
    """forcing_function_second_order(mu) per outer iteration (RIPTRM.py:886-887)."""
    fS = option["forcing_function_second_order"]
    return np.array([fS(float(m)) for m in mu], dtype=np.float64)


def check_supported(option):
    """The GPU path covers both trust-region solvers of the reference: 'tCG' and the class default 'Exact_RepMat' (with or
    without the second-order test); anything else raises -- there is no CPU fallback."""
    if option["TRS_solver"] not in ("tCG", "Exact_RepMat"):
        raise ValueError(f"TRS_solver {option['TRS_solver']} is not supported.")        # RIPTRM.py:453-454
    if option.get("checkTRSoptimality"):
        raise NotImplementedError("checkTRSoptimality is a print-only debug aid (RIPTRM.py:367-388): not on the GPU path")


def to_c_options(option, trace_mode, trace_capacity):
    """Returns (RiptrmOptions, keepalive) for riptrm_set_options."""
    mu, tolL, tolC = barrier_schedule(option)
    o = _lib.RiptrmOptions()
    o.maxiter = int(option["maxiter"])
    o.inner_maxiter = -1 if option["inner_maxiter"] is None else int(option["inner_maxiter"])
    o.tcg_mininner = int(option["tCG_mininner"])
    o.tcg_maxinner = -1
    o.is_euclidean_embedded = int(bool(option["is_euclidean_embedded"]))
    o.trace_mode = int(trace_mode)
    o.trace_capacity = int(trace_capacity)
    o.schedule_split = int(option.get("schedule_split", 0))  # riptrm_b200 extension, see include/riptrm_b200.h
    o.tolresid = float(option["tolresid"])
    o.maxtime = float(option["maxtime"]) if math.isfinite(float(option["maxtime"])) else 1e300
    o.inner_maxtime = -1.0 if option["inner_maxtime"] is None else float(option["inner_maxtime"])
    o.initial_tr_radius = -1.0 if option["initial_TR_radius"] is None else float(option["initial_TR_radius"])
    o.minimal_initial_tr_radius = float(option["minimal_initial_TR_radius"])
    o.maximal_tr_radius = float(option["maximal_TR_radius"])
    o.rho = float(option["rho"])
    o.reduction_regularization = float(option["reduction_regularization"])
    o.gamma = float(option["gamma"])
    o.const_left = float(option["const_left"])
    o.const_right = float(option["const_right"])
    o.tcg_theta = float(option["tCG_theta"])
    o.tcg_kappa = float(option["tCG_kappa"])
    dp = C.POINTER(C.c_double)
    o.mu_sched = mu.ctypes.data_as(dp)
    o.tol_lagrangian_sched = tolL.ctypes.data_as(dp)
    o.tol_complementarity_sched = tolC.ctypes.data_as(dp)
    exact = option["TRS_solver"] == "Exact_RepMat"
    o.trs_solver = _lib.TRS_SOLVER_EXACT_REPMAT if exact else _lib.TRS_SOLVER_TCG
    # the reference evaluates the eigenvalue test only under Exact_RepMat (RIPTRM.py:599): with tCG the flag has no effect
    o.second_order_stationarity = int(bool(option["second_order_stationarity"]) and exact)
    o.trs_tolhardcase = float(option.get("TRS_tolhardcase", 1e-8))
    tolS = second_order_schedule(option, mu) if o.second_order_stationarity else None
    o.tol_second_order_sched = tolS.ctypes.data_as(dp) if tolS is not None else None
    return o, (mu, tolL, tolC, tolS)
