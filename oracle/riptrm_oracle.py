"""TEST INFRASTRUCTURE -- CPU oracle, not product code.

NumPy restatement of the reference's RIPTRM tCG path
(/root/reference/src/solver/RIPTRM.py and the observers in src/solver/utils.py).
It keeps the reference's per-constraint evaluation order (one Riemannian
gradient / Hessian per constraint, accumulated sequentially), so that rounding
follows the reference as closely as a restatement can.

PARITY PIN: the reference has no tests and no golden outputs.
This restatement is pinned against outputs of the UNMODIFIED reference run in this
container on stand-ins for its absent third-party imports
(oracle/run_reference.py, oracle/shims/; fixtures tests/golden/*.json made by
tests/golden/make_golden.py) and against the three notebook known-answers
(SURVEY.md section 4).  The stand-ins restate pymanopt's manifold formulas from its
published source (oracle/manifolds.py) -- that part is "pinned to a restated
dependency", see DESIGN.md.

Differences from the reference, all deliberate:
  * wall-clock stopping (maxtime, inner_maxtime) is not modelled: parity runs are
    iteration-capped; `time` in the log is 0.
  * `tcg_iters` (j+1 of each tCG call, dropped by RIPTRM.py:450) is logged.
  * wandb / verbosity printing are omitted.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import copy
from dataclasses import dataclass, field
from typing import Any

import numpy as np

TCG_STOPS = (
    "MAX_INNER_ITER",
    "NEGATIVE_CURVATURE",
    "EXCEEDED_TR",
    "MODEL_INCREASED",
    "REACHED_TARGET_LINEAR",
    "REACHED_TARGET_SUPERLINEAR",
)


@dataclass
class OracleOutput:
    """utils.Output (src/solver/utils.py:13-16 / base_solver.py:6-11)."""
    name: str
    x: Any
    option: dict
    log: dict
    ineqLagmult: Any
    eqLagmult: Any = field(default_factory=list)


def default_option():
    """RIPTRM.py:305-358 defaults (tCG-path keys only)."""
    return {
        "maxtime": 240,
        "maxiter": 100,
        "tolresid": 1e-15,
        "inner_maxiter": None,
        "inner_maxtime": None,
        "initial_TR_radius": None,
        "minimal_initial_TR_radius": 1e-15,
        "maximal_TR_radius": 10,
        "rho": 0.1,
        "reduction_regularization": 1e3,
        "gamma": 0.25,
        "forcing_function_Lagrangian": lambda mu: max(mu, 1e-14),
        "forcing_function_complementarity": lambda mu: max(1e-3 * mu, 1e-14),
        "min_barrier_parameter": 1e-15,
        "TRS_solver": "tCG",
        "second_order_stationarity": False,
        "do_euclidean_lincomb": False,
        "is_euclidean_embedded": False,
        "forcing_function_second_order": lambda mu: mu,
        "TRS_tolresid": 1e-12,
        "TRS_tolhardcase": 1e-8,
        "basisfun": None,                 # RIPTRM.py:341 default is the RANDOM tangentorthobasis; tests pass a fixed one
        "tCG_theta": 1,
        "tCG_kappa": 0.1,
        "tCG_mininner": 1,
        "initial_barrier_parameter": 0.1,
        "barrier_parameter_update_r": 0.01,
        "barrier_parameter_update_c": 0.5,
        "barrier_parameter_update_b": 0.8,
        "do_simple_barrier_parameter_update": True,
        "const_left": 0.5,
        "const_right": 1e20,
        "manviofun": lambda problem, x: 0,
        "callbackfun": lambda problem, x, y, z, ev: ev,
        "save_inner_iteration": True,
    }


# --------------------------------------------------------------------------
# Steihaug-Toint truncated CG  (RIPTRM.py:41-216, use_rand=False branch)
# --------------------------------------------------------------------------
def steihaug_tcg(man, hess, x, grad, Delta, theta, kappa, mininner, maxinner, precon):
    ip = man.inner_product
    eta = man.zero_vector(x)
    Heta = man.zero_vector(x)          # :47
    r = grad                           # :48
    e_Pe = 0                           # :49
    r_r = ip(x, r, r)                  # :56
    norm_r0 = np.sqrt(r_r)             # :57-58
    z = precon(x, r)                   # :62
    z_r = ip(x, z, r)                  # :67
    d_Pd = z_r                         # :68
    delta = -z                         # :71
    e_Pd = 0                           # :73
    model_value = 0                    # :90
    stop = "MAX_INNER_ITER"            # :95
    j = -1
    for j in range(int(maxinner)):     # :98
        Hdelta = hess(x, delta)        # :100
        d_Hd = ip(x, delta, Hdelta)    # :103
        if d_Hd != 0:                  # :106-114
            alpha = z_r / d_Hd
            e_Pe_new = e_Pe + 2 * alpha * e_Pd + alpha ** 2 * d_Pd
        else:
            e_Pe_new = e_Pe
        if d_Hd <= 0 or e_Pe_new >= Delta ** 2:   # :118
            tau = (-e_Pd + np.sqrt(e_Pd * e_Pd + d_Pd * (Delta ** 2 - e_Pe))) / d_Pd   # :123-125
            eta = eta + tau * delta               # :127
            Heta = Heta + tau * Hdelta            # :132
            stop = "NEGATIVE_CURVATURE" if d_Hd <= 0 else "EXCEEDED_TR"   # :142-145
            break
        e_Pe = e_Pe_new                           # :149
        new_eta = eta + alpha * delta             # :150
        new_Heta = Heta + alpha * Hdelta          # :154
        new_model_value = ip(x, new_eta, grad) + 0.5 * ip(x, new_eta, new_Heta)   # :86-87,162
        if new_model_value >= model_value:        # :163
            stop = "MODEL_INCREASED"
            break
        eta, Heta, model_value = new_eta, new_Heta, new_model_value   # :167-169
        r = r + alpha * Hdelta                    # :172
        r_r = ip(x, r, r)                         # :175
        norm_r = np.sqrt(r_r)
        if j >= mininner and norm_r <= norm_r0 * min(norm_r0 ** theta, kappa):   # :183-185
            stop = "REACHED_TARGET_LINEAR" if kappa < norm_r0 ** theta else "REACHED_TARGET_SUPERLINEAR"
            break
        z = precon(x, r)                          # :195
        zold_rold = z_r                           # :200
        z_r = ip(x, z, r)                         # :202
        beta = z_r / zold_rold                    # :205
        delta = -z + beta * delta                 # :206
        delta = man.to_tangent_space(x, delta)    # :210
        e_Pd = beta * (e_Pd + alpha * d_Pd)       # :213
        d_Pd = z_r + beta * beta * d_Pd           # :214
    return eta, Heta, j, stop


# --------------------------------------------------------------------------
# Lagrangian / barrier operators  (RIPTRM.py:457-571)
# --------------------------------------------------------------------------
def _is_product(man):
    return hasattr(man, "manifolds")


def _amb(man, v):
    """RIPTRM.py:13-38 `_ProductAmbientVector`: list algebra for ambient product vectors."""
    if _is_product(man):
        from .manifolds import _TangentList
        return _TangentList(v)
    return v


def egrad_lagrangian(problem, x, y):
    """RIPTRM.py:457-473."""
    man = problem.manifold
    vec = _amb(man, problem.euclidean_gradient(x))
    negs = [-_amb(man, eg(x)) for eg in problem.ineqconstraints_euclidean_gradient_all]
    for i in range(len(y)):
        vec = vec - y[i] * negs[i]
    return vec


def grad_lagrangian(problem, x, y, lincomb=False):
    """RIPTRM.py:475-489."""
    if getattr(problem, "closed", None) is not None:
        return problem.closed.grad_lagrangian(x, y)
    man = problem.manifold
    if lincomb:
        return man.euclidean_to_riemannian_gradient(x, egrad_lagrangian(problem, x, y))
    vec = problem.riemannian_gradient(x)
    negs = [-g(x) for g in problem.ineqconstraints_riemannian_gradient_all]
    for i in range(len(y)):
        vec = vec - y[i] * negs[i]
    return vec


def hess_lagrangian(problem, x, y, dx, lincomb=False):
    """RIPTRM.py:491-523."""
    if getattr(problem, "closed", None) is not None:
        return problem.closed.hess_lagrangian(x, y, dx)
    man = problem.manifold
    if lincomb:
        eg = egrad_lagrangian(problem, x, y)
        vec = _amb(man, problem.euclidean_hessian(x, dx))
        negs = [-_amb(man, eh(x, dx)) for eh in problem.ineqconstraints_euclidean_hessian_all]
        for i in range(len(y)):
            vec = vec - y[i] * negs[i]
        return man.euclidean_to_riemannian_hessian(x, eg, vec, dx)
    vec = problem.riemannian_hessian(x, dx)
    negs = [-h(x, dx) for h in problem.ineqconstraints_riemannian_hessian_all]
    for i in range(len(y)):
        vec = vec - y[i] * negs[i]
    return vec


def G_apply(problem, x, w, lincomb=False):
    """G_x(w) = sum_i w_i grad s_i(x), s_i = -g_i.  RIPTRM.py:525-551."""
    if getattr(problem, "closed", None) is not None:
        return problem.closed.G_apply(x, w)
    man = problem.manifold
    if lincomb:
        negs = [-_amb(man, eg(x)) for eg in problem.ineqconstraints_euclidean_gradient_all]
        vec = man.zero_vector(x)
        for i in range(len(negs)):
            vec = vec + w[i] * negs[i]
        return man.euclidean_to_riemannian_gradient(x, vec)
    negs = [-g(x) for g in problem.ineqconstraints_riemannian_gradient_all]
    vec = man.zero_vector(x)
    for i in range(len(negs)):
        vec = vec + w[i] * negs[i]
    return vec


def Gadj_apply(problem, x, dx, euclidean_embedded=False):
    """G*_x[dx]_i = <grad s_i(x), dx>_x.  RIPTRM.py:553-571."""
    if getattr(problem, "closed", None) is not None:
        return problem.closed.Gadj_apply(x, dx, euclidean_embedded)
    man = problem.manifold
    if euclidean_embedded:
        negs = [-_amb(man, eg(x)) for eg in problem.ineqconstraints_euclidean_gradient_all]
    else:
        negs = [-g(x) for g in problem.ineqconstraints_riemannian_gradient_all]
    return np.array([man.inner_product(x, g, dx) for g in negs])


def slack(problem, x):
    """costineqconstvecfun: s(x) = -g(x).  RIPTRM.py:576,721."""
    if getattr(problem, "closed", None) is not None:
        return problem.closed.slack(x)
    return np.array([-g(x) for g in problem.ineqconstraints_all])


# --------------------------------------------------------------------------
# Observers  (utils.py:237-368)
# --------------------------------------------------------------------------
def kkt_residual(problem, x, y, manviofun):
    """utils.compute_residual (utils.py:269-340), inequality-only."""
    man = problem.manifold
    closed = getattr(problem, "closed", None)
    if closed is not None:
        vec = closed.grad_lagrangian(x, y)
        gvals = list(-closed.slack(x))
    else:
        vec = problem.riemannian_gradient(x)
        grads = problem.ineqconstraints_riemannian_gradient_all
        for i in range(problem.num_ineqconstraints):
            vec = vec + y[i] * grads[i](x)
        gvals = [g(x) for g in problem.ineqconstraints_all]
    gradnorm = man.norm(x, vec)
    sq_compl = 0
    for i, gv in enumerate(gvals):
        sq_compl += (y[i] * gv) ** 2
    sq_nonneg = 0
    for yv in y:
        sq_nonneg += max(-yv, 0) ** 2
    sq_ineq = 0
    for gv in gvals:
        sq_ineq += max(gv, 0) ** 2
    manvio = manviofun(problem, x)
    residual = np.sqrt(gradnorm ** 2 + sq_compl + sq_nonneg + sq_ineq + 0 + manvio ** 2)
    return residual, gradnorm, np.sqrt(sq_compl), np.sqrt(sq_nonneg), manvio


def evaluate(problem, xPrev, x, y, manviofun, callbackfun):
    """utils.evaluation (utils.py:342-368) + compute_maxmeanviolations (:237-267)."""
    cost = problem.cost(x)
    try:
        dist = problem.manifold.dist(xPrev, x)
    except NotImplementedError:   # pymanopt's Stiefel: the reference's evaluation would stop here; the log gets NaN
        dist = float("nan")
    residual, gradnorm, compl, nonneg, manvio = kkt_residual(problem, x, y, manviofun)
    maxv = 0
    meanv = 0
    closed = getattr(problem, "closed", None)
    gvals = list(-closed.slack(x)) if closed is not None else [g(x) for g in problem.ineqconstraints_all]
    for gv in gvals:
        v = max(gv, 0)
        maxv = max(maxv, v)
        meanv += v
    if problem.num_ineqconstraints > 0:
        meanv = meanv / problem.num_ineqconstraints
    ev = {
        "cost": cost, "distance": dist, "residual": residual, "gradnorm": gradnorm,
        "complviolation": compl, "dualviolation": nonneg, "manviolation": manvio,
        "maxviolation": maxv, "meanviolation": meanv,
    }
    return callbackfun(problem, x, y, [], ev)


# --------------------------------------------------------------------------
# Exact trust-region subproblem on the representation matrix (TRS_solver='Exact_RepMat')
# --------------------------------------------------------------------------
def operator_matrix(man, x, F, basis):
    """selfadj_operator2matrix (utils.py:565-573): upper triangle from <F(b_j), b_i>, mirrored."""
    n = len(basis)
    A = np.zeros((n, n))
    for j in range(n):
        Fb = F(basis[j])
        for i in range(j + 1):
            A[i, j] = man.inner_product(x, Fb, basis[i])
    return A + np.triu(A, 1).T


def tangent_coords(man, x, basis, v):
    """The loops of RIPTRM.py:436-438 / :605-607: coefficient i = <v, b_i>_x."""
    out = np.empty(len(basis))
    for i in range(len(basis)):
        out[i] = man.inner_product(x, v, basis[i])
    return out


def trs_gep(A, a, B, Del, tolhardcase=1e-4):
    """min x'Ax/2 + a'x s.t. x'Bx <= Del^2 through the rightmost eigenpair of a 2n x 2n pencil
    (RIPTRM.py:218-299; Adachi, Iwata, Nakatsukasa, Takeda 2017).  Same library calls as the reference:
    scipy.sparse.linalg.cg for the interior candidate (:244), scipy.linalg.eig for the pencil (:252),
    scipy.linalg.solve / eigh in the hard case (:270-280)."""
    import scipy.linalg
    import scipy.sparse.linalg
    n = A.shape[0]
    aat = np.outer(a, a) / (Del ** 2)
    M0 = np.block([[-B, A], [A, -aat]])                                  # :240
    M1 = np.block([[np.zeros((n, n)), B], [B, np.zeros((n, n))]])        # :241
    newton, _ = scipy.sparse.linalg.cg(A, -a)                            # :244 (rtol 1e-5, maxiter 10 n)
    ok = np.linalg.norm(A @ newton + a) / np.linalg.norm(a) < 1e-5       # :245
    if ok and newton @ B @ newton >= Del ** 2:                           # :246-247
        ok = False
    if not ok:
        newton = np.full_like(newton, np.nan)                            # :247, :249
    lams, vecs = scipy.linalg.eig(a=M0, b=-M1)                           # :252
    k = np.argmax(np.real(lams))                                         # :253
    lam1 = np.real(lams[k])
    V = np.real(vecs[:, k])                                              # :255-256
    x = V[:n]                                                            # :257
    normx = np.sqrt(x @ (B @ x))                                         # :258
    x = x / normx * Del                                                  # :259
    if x @ a > 0:                                                        # :260-261
        x = -x
    kind = "boundary"
    if normx < tolhardcase:                                              # :263
        x1 = V[n:]
        shifted = A + lam1 * B                                           # :267
        Bp = B @ x1
        H = shifted + lam1 * np.outer(Bp, Bp)                            # :268-269 (alpha1 = lam1)
        x2 = scipy.linalg.solve(H, -a, assume_a="sym")                   # :270
        kind = "hardcase_1"
        if np.linalg.norm(shifted @ x2 + a) / np.linalg.norm(a) > tolhardcase:   # :274
            _, v = scipy.linalg.eigh(A, B)                               # :275
            for ii in (3, 6, 9):                                         # :276-283
                Bp = B @ v[:, :ii]
                H = shifted + lam1 * Bp @ Bp.T
                x2 = scipy.linalg.solve(H, -a, assume_a="sym")
                kind = f"hardcase_{ii}"
                if np.linalg.norm(shifted @ x2 + a) / np.linalg.norm(a) < tolhardcase:
                    break
        Bx, Bx2 = B @ x1, B @ x2                                         # :284-285
        aa, bb, cc = x1 @ Bx, 2 * x2 @ Bx, x2 @ Bx2 - Del ** 2           # :286-288
        alp = (-bb + np.sqrt(bb ** 2 - 4 * aa * cc)) / (2 * aa)          # :289
        x = x2 + alp * x1                                                # :290
    if not np.isnan(newton).any():                                       # :293-299
        if 0.5 * (newton @ A @ newton) + a @ newton <= 0.5 * (x @ A @ x) + a @ x:
            return newton, 0, "interior"
    return x, lam1, kind


# --------------------------------------------------------------------------
# The condensed Newton system of the reference's interior-point method (RIPM.py:484-511), same operator as Hw
# --------------------------------------------------------------------------
def newton_operator(problem, x, z, s, lincomb=False):
    """OperatorAw = OperatorHessLag + OperatorTHETA (RIPM.py:491-493): dx -> Hess L(x, z)[dx] + G_x(G*_x[dx] * z / s), the
    slack s an independent variable."""
    def Aw(dx):
        return hess_lagrangian(problem, x, z, dx, lincomb) + G_apply(problem, x, Gadj_apply(problem, x, dx) * (z / s), lincomb)
    return Aw


def newton_repmat(problem, x, z, s, c, basis):
    """RepresentMatMethod without equality constraints (RIPM.py:238-300): T_mat = Aw_mat, scipy.linalg.solve(assume_a='sym')."""
    import scipy.linalg
    man = problem.manifold
    Aw = newton_operator(problem, x, z, s)
    Amat = operator_matrix(man, x, Aw, basis)
    sol = scipy.linalg.solve(Amat, tangent_coords(man, x, basis, c), assume_a="sym")
    dx = man.zero_vector(x)
    for i in range(len(basis)):
        dx = dx + sol[i] * basis[i]
    return dx, Amat


def conj_res(man, x, A, b, tol, maxiter):
    """TangentSpaceConjResMethod (utils.py:582-618; Saad, Iterative Methods for Sparse Linear Systems, Alg. 6.20), v0 = 0."""
    v = man.zero_vector(x)
    r = b
    p = copy.deepcopy(r)
    b_norm = man.norm(x, b)
    Ar = A(r)
    Ap = A(p)
    rAr = man.inner_product(x, r, Ar)
    t = 0
    while True:
        t += 1
        a = rAr / man.inner_product(x, Ap, Ap)
        v = v + a * p
        r = r - a * Ap
        rel_res = man.norm(x, r) / b_norm
        if rel_res < tol or t == maxiter:
            break
        Ar = A(r)
        old = rAr
        rAr = man.inner_product(x, r, Ar)
        beta = rAr / old
        p = r + beta * p
        Ap = Ar + beta * Ap
    return v, t, rel_res


# --------------------------------------------------------------------------
# The solver
# --------------------------------------------------------------------------
class OracleRIPTRM:
    def __init__(self, option=None):
        opt = default_option()
        opt.update(option or {})
        if opt["TRS_solver"] not in ("tCG", "Exact_RepMat"):
            raise ValueError(f"TRS_solver {opt['TRS_solver']} is not supported.")        # RIPTRM.py:453-454
        if opt["TRS_solver"] == "Exact_RepMat" and opt["basisfun"] is None:
            raise ValueError("Exact_RepMat: pass option['basisfun'] (the reference's default basis is random)")
        self.rep = None      # (basis, Hw matrix, c vector) at the current (x, y): is_RepMat_available (RIPTRM.py:415-421)
        self.rep_new = None
        self.option = opt
        self.log = {}
        self.name = f"RIPTRM_{opt['TRS_solver']}"
        self.counters = {"tcg_hessvec": 0, "aux_hessvec": 0, "inner": 0, "tcg_calls": 0}

    # base_solver.py:58-76
    def _add_log(self, it, ev, status):
        row = {"iteration": it, "time": 0.0}
        row.update(ev)
        row.update(status)
        if it == 0 and not self.log:
            for k, v in row.items():
                self.log[k] = [v]
        else:
            for k, v in row.items():
                self.log[k].append(v)

    # RIPTRM.py:980-1024 (+ tcg_iters)
    def _status(self, y, mu, info):
        keys = ("num_inner", "inner_status", "TR_radius", "dxtype", "normdx", "minxfeasi", "minyfeasi",
                "compl", "mineigvalHw", "ared/pred", "radius_update", "dual_clipping", "tcg_iters")
        st = {"mu": mu}
        for k in keys:
            st[k] = None if info is None else info.get(k)
        m = float("-inf")
        for v in y:
            m = max(m, abs(v))
        st["maxabsLagmult"] = m
        return st

    # RIPTRM.py:574-629
    def _inner_stop_tests(self, problem, xNew, yNew, mu, tolL, tolC, tolS=None):
        o = self.option
        man = problem.manifold
        lin, emb = o["do_euclidean_lincomb"], o["is_euclidean_embedded"]
        sNew = slack(problem, xNew)
        ngl = man.norm(xNew, grad_lagrangian(problem, xNew, yNew, lin))
        compl = np.linalg.norm(yNew * sNew - mu)
        mineig, eig_ok = None, True
        if o["TRS_solver"] == "Exact_RepMat" and o["second_order_stationarity"]:     # :599-617
            import scipy.linalg

            def HwNew(dx):
                return hess_lagrangian(problem, xNew, yNew, dx, lin) + G_apply(
                    problem, xNew, (yNew * Gadj_apply(problem, xNew, dx, emb)) / sNew, lin)

            basis = o["basisfun"](man, xNew)
            Hmat = operator_matrix(man, xNew, HwNew, basis)
            cNew = problem.riemannian_gradient(xNew) - G_apply(problem, xNew, mu / sNew, lin)
            cvec = tangent_coords(man, xNew, basis, cNew)
            mineig = scipy.linalg.eigh(Hmat, eigvals_only=True)[0]                   # :609-610
            eig_ok = bool(mineig >= -tolS)                                           # :611
            self.rep_new = (basis, Hmat, cvec)                                       # :613-615
        return {
            "xfeasi": bool(np.all(sNew > 0)),
            "yfeasi": bool(np.all(yNew > 0)),
            "gradL": bool(ngl <= tolL),
            "compl_ok": bool(compl <= tolC),
            "eig_ok": eig_ok, "mineigvalHw": mineig,
            "minxfeasi": min(sNew), "minyfeasi": min(yNew), "compl": compl, "sNew": sNew,
        }

    # RIPTRM.py:631-705
    def _rho_test_and_update(self, problem, x, y, Hw, c, dx, normdx, xNew, yNew, sNew, mu, Delta):
        o = self.option
        man = problem.manifold
        s = slack(problem, x)

        def logbarr(pt, sv):              # :644-649 (recomputes the cost: cached value ignored)
            return problem.cost(pt) - mu * np.sum(np.log(sv))

        phi_cur = logbarr(x, s)
        phi_new = logbarr(xNew, sNew)
        ared = phi_cur - phi_new                                                     # :658
        self.counters["aux_hessvec"] += 1
        pred = 0 - 0.5 * man.inner_product(x, Hw(dx), dx) - man.inner_product(x, c, dx)   # :659
        reg = max(1, abs(phi_cur)) * np.spacing(1) * o["reduction_regularization"]   # :660
        ared = ared + reg
        pred = pred + reg
        out = {"ared/pred": ared / pred}
        if ared < 0.25 * pred:                                                       # :667-675
            out["radius_update"] = "reduced"
            DeltaNext = 0.25 * Delta
        elif ared >= 0.75 * pred and np.abs(normdx - Delta) <= 1e-15:
            out["radius_update"] = "expanded"
            DeltaNext = min(2 * Delta, o["maximal_TR_radius"])
        else:
            out["radius_update"] = "unchanged"
            DeltaNext = Delta
        if ared > o["rho"] * pred:                                                   # :677
            out["inner_status"] = "successful"
            I_left = o["const_left"] * np.minimum(np.minimum(y, mu / sNew), 1)       # :681
            # :682 -- np.maximum(a, b, out): the third positional argument is `out`, so the
            # y / s terms are silently dropped: I_right = max(const_right, const_right/mu).
            I_right = np.full_like(y, max(o["const_right"], o["const_right"] / mu))
            clipped = np.minimum(np.maximum(yNew, I_left), I_right)                  # :683-684
            out["dual_clipping"] = not np.array_equal(yNew, clipped)                 # :685-695
            # :687-695: the matrix built at (xNew, yNew) for the eigenvalue test is next iteration's matrix, unless
            # clipping changed the multipliers
            self.rep = None
            if (not out["dual_clipping"] and o["TRS_solver"] == "Exact_RepMat" and o["second_order_stationarity"]):
                self.rep = copy.deepcopy(self.rep_new)
            return copy.deepcopy(xNew), clipped, DeltaNext, out
        out["inner_status"] = "unsuccessful"                                         # :697-702 (matrix stays valid)
        out["dual_clipping"] = None
        return x, y, DeltaNext, out

    # RIPTRM.py:707-783
    def _inner_step(self, problem, x, y, mu, Delta, k_inner, tolL, tolC, tolS=None):
        o = self.option
        man = problem.manifold
        lin, emb = o["do_euclidean_lincomb"], o["is_euclidean_embedded"]
        info = {"num_inner": k_inner, "TR_radius": Delta, "mineigvalHw": None,
                "ared/pred": None, "radius_update": None, "dual_clipping": None}
        s = slack(problem, x)                                                        # :725
        gradf = problem.riemannian_gradient(x)                                       # :726

        def Hw(dx):                                                                  # :729
            return hess_lagrangian(problem, x, y, dx, lin) + G_apply(
                problem, x, (y * Gadj_apply(problem, x, dx, emb)) / s, lin)

        c = gradf - G_apply(problem, x, mu / s, lin)                                 # :730

        def counted_Hw(_x, dx):
            self.counters["tcg_hessvec"] += 1
            return Hw(dx)

        if o["TRS_solver"] == "Exact_RepMat":                                        # :431-444
            if self.rep is None:
                basis = o["basisfun"](man, x)
                self.rep = (basis, operator_matrix(man, x, Hw, basis), tangent_coords(man, x, basis, c))
            basis, Hmat, cvec = self.rep
            coeff, _lam1, kind = trs_gep(Hmat, cvec, np.eye(man.dim), Delta, o["TRS_tolhardcase"])
            dx = man.zero_vector(x)
            for i in range(man.dim):
                dx = dx + coeff[i] * basis[i]
            info["dxtype"] = kind
            info["tcg_iters"] = None
        else:
            dx, _Heta, j, stop = steihaug_tcg(                                       # :445-452
                man, counted_Hw, x, c, Delta, o["tCG_theta"], o["tCG_kappa"], o["tCG_mininner"],
                man.dim, problem.preconditioner)
            self.counters["tcg_calls"] += 1
            info["dxtype"] = f"tCG_{stop}"
            info["tcg_iters"] = j + 1
        normdx = man.norm(x, dx)                                                     # :735
        info["normdx"] = normdx
        dy = -y + mu * (1 / s) - y * Gadj_apply(problem, x, dx, emb) / s             # :743
        xNew = man.retraction(x, dx)                                                 # :744
        yNew = y + dy                                                                # :745
        t = self._inner_stop_tests(problem, xNew, yNew, mu, tolL, tolC, tolS)        # :748
        info.update(minxfeasi=t["minxfeasi"], minyfeasi=t["minyfeasi"], compl=t["compl"], mineigvalHw=t["mineigvalHw"])
        if t["xfeasi"] and t["yfeasi"] and t["gradL"] and t["compl_ok"] and t["eig_ok"]:   # :762-766
            info["inner_status"] = "converged"
            return True, xNew, yNew, Delta, info
        if not t["xfeasi"]:                                                          # :769-775 (matrix stays valid)
            info["inner_status"] = "primal_infeasible"
            return False, x, y, o["gamma"] * normdx, info
        xN, yN, DeltaN, upd = self._rho_test_and_update(                             # :777
            problem, x, y, Hw, c, dx, normdx, xNew, yNew, t["sNew"], mu, Delta)
        info.update(upd)
        return False, xN, yN, DeltaN, info

    # RIPTRM.py:785-847 (iteration guards only)
    def _inner_run(self, problem, k_outer, x0, y0, mu, Delta0, tolL, tolC, tolS=None):
        o = self.option
        x, y, Delta = x0, y0, Delta0
        self.rep = self.rep_new = None                                               # inner_preprocess :415-421
        xPrev = copy.deepcopy(x)
        k = 0
        while True:
            k += 1
            done, x, y, Delta, info = self._inner_step(problem, x, y, mu, Delta, k, tolL, tolC, tolS)
            self.counters["inner"] += 1
            if o["save_inner_iteration"]:                                            # :812-818
                ev = evaluate(problem, xPrev, x, y, o["manviofun"], o["callbackfun"])
                self._add_log(k_outer, ev, self._status(y, mu, info))
            xPrev = copy.deepcopy(x)
            if o["inner_maxiter"] is not None and k >= o["inner_maxiter"]:           # :835-842
                info["inner_status"] = "max-iter-exceeded"
                done = True
                x, y, Delta = x0, y0, Delta0
            if done:
                break
        return x, y, Delta, info

    # RIPTRM.py:909-976
    def run(self, problem):
        o = self.option
        x = copy.deepcopy(problem.initialpoint)                                      # :849-864
        y = copy.deepcopy(problem.initialineqLagmult)
        mu = o["initial_barrier_parameter"]
        if o["initial_TR_radius"] is None:
            Delta = problem.manifold.typical_dist / 8
        else:
            Delta = o["initial_TR_radius"]
        xPrev = copy.deepcopy(x)
        info = None
        it = 0
        while True:
            ev = evaluate(problem, xPrev, x, y, o["manviofun"], o["callbackfun"])    # :933
            if it == 0 or not o["save_inner_iteration"]:                             # :936-941
                self._add_log(it, ev, self._status(y, mu, info))
            xPrev = copy.deepcopy(x)
            # base_solver.py:85-106 (time criterion omitted)
            reason = None
            if it >= o["maxiter"]:
                reason = f"Max iteration count reached; maxiter={o['maxiter']}"
            if ev["residual"] <= o["tolresid"]:
                reason = ("KKT residual tolerance reached; current residual=" + str(ev["residual"])
                          + " and tolresid=" + str(o["tolresid"]))
            if reason is not None:
                o["stoppingcriterion"] = reason
                break
            it += 1
            # outer_step, RIPTRM.py:866-896
            tolL = o["forcing_function_Lagrangian"](mu)
            tolC = o["forcing_function_complementarity"](mu)
            tolS = o["forcing_function_second_order"](mu) if o["second_order_stationarity"] else None
            x, y, Delta, info = self._inner_run(problem, it, x, y, mu, Delta, tolL, tolC, tolS)
            r_, c_, b_ = (o["barrier_parameter_update_r"], o["barrier_parameter_update_c"],
                          o["barrier_parameter_update_b"])
            if o["do_simple_barrier_parameter_update"]:                              # :890-893
                mu = max(o["min_barrier_parameter"], c_ * (mu ** (1 + r_)))
            else:
                mu = max(o["min_barrier_parameter"], min(b_ * mu, c_ * (mu ** (1 + r_))))
            Delta = max(Delta, o["minimal_initial_TR_radius"])                       # :894
        return OracleOutput(name=self.name, x=x, ineqLagmult=y, eqLagmult=[],
                            option=copy.deepcopy(o), log=self.log)
