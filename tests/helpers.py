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


# ----------------------------------------------------------------------------------------------------------------
# NumPy restatement of the device-side generator (csrc/datagen.cuh): Philox4x32-10, polar normals, det_log
# ----------------------------------------------------------------------------------------------------------------
def _philox(key, index, stream, sub):
    """Vectorised over `index` (uint64 array); key, stream, sub scalars or arrays.  Returns four uint32 arrays."""
    index = np.asarray(index, dtype=np.uint64)
    M = np.uint64(0xFFFFFFFF)
    c0 = index & M
    c1 = index >> np.uint64(32)
    c2 = np.broadcast_to(np.uint64(stream), index.shape).copy()
    c3 = np.broadcast_to(np.asarray(sub, dtype=np.uint64), index.shape).copy()
    k0 = np.uint64(int(key) & 0xFFFFFFFF)
    k1 = np.uint64((int(key) >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n1 = p1 & M
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        n3 = p0 & M
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & M
        k1 = (k1 + np.uint64(0xBB67AE85)) & M
    return c0, c1, c2, c3


def _u53(hi, lo):
    return ((hi >> np.uint64(5)).astype(np.float64) * 67108864.0 + (lo >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def det_log_np(x):
    """csrc/common.cuh det_log (fdlibm-style, no contraction) for positive normal doubles."""
    x = np.asarray(x, dtype=np.float64)
    u = x.view(np.uint64).copy()
    hx = u >> np.uint64(32)
    k = (hx >> np.uint64(20)).astype(np.int64) - 1023
    hx = hx & np.uint64(0x000FFFFF)
    i = (hx + np.uint64(0x95F64)) & np.uint64(0x100000)
    u = ((hx | (i ^ np.uint64(0x3FF00000))) << np.uint64(32)) | (u & np.uint64(0xFFFFFFFF))
    k = k + (i >> np.uint64(20)).astype(np.int64)
    f = u.view(np.float64) - 1.0
    dk = k.astype(np.float64)
    Lg = (6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
          1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01)
    ln2_hi, ln2_lo = 6.93147180369123816490e-01, 1.90821492927058770002e-10
    s = f / (2.0 + f)
    z = s * s
    w = z * z
    t1 = w * (Lg[1] + w * (Lg[3] + w * Lg[5]))
    t2 = z * (Lg[0] + w * (Lg[2] + w * (Lg[4] + w * Lg[6])))
    R = t2 + t1
    hfsq = 0.5 * f * f
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f)


def _normal_twin(key, index, stream):
    index = np.asarray(index, dtype=np.uint64)
    out = np.zeros(index.shape)
    todo = np.ones(index.shape, dtype=bool)
    for t in range(64):
        if not todo.any():
            break
        a, b, c, d = _philox(key, index[todo], stream, t)
        v1, v2 = 2.0 * _u53(a, b) - 1.0, 2.0 * _u53(c, d) - 1.0
        s = v1 * v1 + v2 * v2
        ok = (s < 1.0) & (s > 0.0)
        val = np.zeros(s.shape)
        val[ok] = v1[ok] * np.sqrt((-2.0 * det_log_np(s[ok])) / s[ok])
        idx = np.flatnonzero(todo)
        out[idx[ok]] = val[ok]
        todo[idx[ok]] = False
    return out


def device_generator_twin(inst, n, points, snr=0.5, delta=0.7):
    """(Z [n, n], x0 [points, n]) of instance `inst` exactly as csrc/datagen.cuh draws them."""
    k = int(np.floor(delta * n))
    perm = np.arange(n)
    a, b, _, _ = _philox(inst, np.arange(k, dtype=np.uint64), 0, 0)
    us = _u53(a, b)
    for i in range(k):
        j = min(i + int(us[i] * float(n - i)), n - 1)
        perm[i], perm[j] = perm[j], perm[i]
    v = np.zeros(n)
    v[perm[:k]] = 1.0 / np.sqrt(float(k))
    rn = np.sqrt(float(n))
    noise = _normal_twin(inst, np.arange(n * n, dtype=np.uint64), 1).reshape(n, n) / rn
    diag = (_normal_twin(inst, np.arange(n, dtype=np.uint64), 2) * 2.0) / rn
    noise[np.arange(n), np.arange(n)] = diag
    Z = np.sqrt(snr) * np.outer(v, v) + noise
    x0 = np.empty((points, n))
    for pt in range(points):
        a, b, _, _ = _philox(inst, np.arange(n, dtype=np.uint64), 3 + pt, 0)
        u = _u53(a, b)
        s = 0.0
        for val in u:
            s = s + val * val
        x0[pt] = np.abs(u / np.sqrt(s))
    return Z, x0


# ----------------------------------------------------------------------------------------------------------------
# per-outer-iteration view of a log (profiles/parity_r02.md, tests/test_parity_report.py)
# ----------------------------------------------------------------------------------------------------------------
def per_outer(log, tcg=None):
    """Aggregates a per-inner-iteration log (SURVEY.md App. E layout: row 0 + one row per trust-region iteration,
    `iteration` = outer index) into per-outer-iteration arrays: trust-region iterations, tCG iterations, and the
    last row's radius / cost / residual / status.  `tcg`: per-row tCG counts (`log['tcg_iters']` when absent)."""
    it = np.array(log["iteration"], dtype=int)
    if tcg is None:
        tcg = log["tcg_iters"]
    tcg = np.array([0 if v is None else v for v in tcg], dtype=float)
    if len(tcg) == len(it) - 1:          # golden files keep one count per tCG call (no entry for row 0)
        tcg = np.concatenate([[0.0], tcg])
    K = int(it.max())
    out = {k: [] for k in ("outer", "inner", "tcg", "radius", "cost", "residual", "mu", "status", "last_row")}
    for k in range(1, K + 1):
        rows = np.nonzero(it == k)[0]
        rows = rows[rows > 0]
        if len(rows) == 0:
            continue
        last = int(rows[-1])
        out["outer"].append(k)
        out["inner"].append(len(rows))
        out["tcg"].append(float(tcg[rows].sum()))
        out["radius"].append(float(log["TR_radius"][last]))
        out["cost"].append(float(log["cost"][last]))
        out["residual"].append(float(log["residual"][last]))
        out["mu"].append(float(log["mu"][last]))
        out["status"].append(log["inner_status"][last])
        out["last_row"].append(last)
    return {k: (np.array(v) if k not in ("status",) else v) for k, v in out.items()}


def outer_window(a, b, key):
    """Number of leading outer iterations for which per-outer arrays a[key] and b[key] are identical."""
    n = min(len(a[key]), len(b[key]))
    for i in range(n):
        if a[key][i] != b[key][i]:
            return i
    return n
