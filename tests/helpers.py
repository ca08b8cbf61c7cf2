"""Shared helpers for the parity tests: building oracle problems from the fixtures and
comparing solver logs column by column."""
import numpy as np

from oracle.problems import (NonnegPCAProblem, RosenbrockProblem,
                             StableIdentificationProblem)

FLOAT_COLUMNS = ("cost", "residual", "gradnorm", "complviolation", "dualviolation", "manviolation",
                 "maxviolation", "meanviolation", "mu", "TR_radius", "normdx", "minxfeasi", "minyfeasi",
                 "compl", "ared/pred", "maxabsLagmult", "distance")
DISCRETE_COLUMNS = ("iteration", "num_inner", "inner_status", "dxtype", "radius_update", "dual_clipping")


def nonnegpca_problem(datasets):
    d = datasets["NonnegPCA/1"]
    return NonnegPCAProblem(d["Z"], d["initx_a"], d["initineqLagmult"])


def stableid_problem(datasets, pt):
    d = datasets["StableIdentification/1"]
    Xs = [d[f"noisyX_{k}"] for k in range(1, 6)]
    X = np.hstack([x[:, :-1] for x in Xs])
    XP = np.hstack([x[:, 1:] for x in Xs])
    x0 = [d[f"init{c}_{pt}"] for c in "JRQ"]
    return StableIdentificationProblem(X, XP, 0.02, d["constset"], x0, d["initineqLagmult"])


def rosenbrock_problem():
    return RosenbrockProblem(5, 3, 1e7)


def _num(col):
    return np.array([np.nan if v is None else float(v) for v in col], dtype=float)


def first_discrete_mismatch(log_a, log_b, columns=DISCRETE_COLUMNS):
    """Index of the first row where any discrete column differs (len if none)."""
    n = min(len(log_a["iteration"]), len(log_b["iteration"]))
    for i in range(n):
        for k in columns:
            if k in log_a and k in log_b and log_a[k][i] != log_b[k][i]:
                return i
    return n if len(log_a["iteration"]) == len(log_b["iteration"]) else n


def max_rel_diff(log_a, log_b, column, rows=None, floor=0.0):
    a, b = _num(log_a[column]), _num(log_b[column])
    n = min(len(a), len(b)) if rows is None else rows
    a, b = a[:n], b[:n]
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    mask = ~(np.isnan(a) & np.isnan(b)) & ~both_inf
    if not mask.any():
        return 0.0
    den = np.maximum(np.abs(b[mask]), floor if floor > 0 else 1e-300)
    return float(np.max(np.abs(a[mask] - b[mask]) / den))


def stiefel_start(n, p, seed):
    """Feasible starting point of the Stiefel reading of config 4: nonnegative orthonormal
    columns with disjoint supports (blocks of n // p rows, |rand| + 0.1, column-normalised)."""
    rs = np.random.RandomState(seed)
    X = np.zeros((n, p))
    b = n // p
    for c in range(p):
        lo, hi = c * b, (n if c == p - 1 else (c + 1) * b)
        X[lo:hi, c] = np.abs(rs.rand(hi - lo)) + 0.1
        X[:, c] /= np.linalg.norm(X[:, c])
    return X
