"""The deterministic C oracle (oracle/c/riptrm_det.c) against the unmodified reference's golden run and the
NumPy oracle.  Tier T2 of SURVEY.md App. C: identical discrete traces inside the well-conditioned window,
objective / iterates to 1e-8 at converged points; tier T1 (bit-for-bit GPU == C oracle) is in the GPU tests."""
import numpy as np
import pytest

import riptrm_b200 as rb
from conftest import load_golden
from helpers import DISCRETE_COLUMNS, first_discrete_mismatch, max_rel_diff
from oracle.c import binding as detc
from oracle.problems import NonnegPCAProblem, nonnegpca_generate_instance
from oracle.riptrm_oracle import OracleRIPTRM

REL_TOL = 1e-8


def _converged_costs(log):
    return np.array([c for c, s in zip(log["cost"], log["inner_status"]) if s == "converged"])


def test_c_oracle_vs_reference_golden(datasets):
    g = load_golden("nonnegpca_1_a_K40")
    G = dict(g["log"], tcg_iters=[None] + g["tcg_iters"])
    d = datasets["NonnegPCA/1"]
    x, y, sm, tr = detc.solve(d["Z"], d["initx_a"], d["initineqLagmult"], {"maxiter": 40, "tolresid": 0},
                              trace_capacity=512)
    L = rb.trace_to_log(tr)
    nrows = len(G["iteration"])
    outer_of = lambda row: G["iteration"][min(row, nrows - 1)]
    first = first_discrete_mismatch(L, G)
    assert outer_of(first) > 16, (first, outer_of(first))          # SURVEY App. C window
    first_tcg = first_discrete_mismatch(L, G, columns=DISCRETE_COLUMNS + ("tcg_iters",))
    assert outer_of(first_tcg) > 8, (first_tcg, outer_of(first_tcg))
    assert max_rel_diff(L, G, "TR_radius", rows=first_tcg) == 0.0   # radii are exact until a tCG count differs
    assert max_rel_diff(L, G, "cost", rows=first) < 1e-6           # transient inner iterates
    a, b = _converged_costs(L), _converged_costs(G)
    assert len(a) == len(b) == 40
    assert np.max(np.abs(a - b) / np.abs(b)) < REL_TOL             # every outer iteration's converged point
    assert np.max(np.abs(x - np.array(g["x"]))) < REL_TOL
    assert np.max(np.abs(y - np.array(g["ineqLagmult"]))) < 1e-6 * max(1.0, np.max(np.abs(g["ineqLagmult"])))
    assert sm[10] == 40 and sm[14] == 2                             # 40 outer iterations, stopped by maxiter
    assert abs(L["residual"][0] - 4.986888432851818) < 1e-13       # notebook known answer


def test_c_oracle_vs_numpy_oracle_generated():
    """A bench-workload pair (generator law, seed 3), the bench protocol (30 outer iterations): identical discrete
    trace while the problem is well conditioned, every outer iteration's converged objective to 1e-6 (inner
    solves are inexact: tolerance mu), and the FINAL objective / iterate / multipliers to 1e-8 (north_star)."""
    Z, x0, y0 = nonnegpca_generate_instance(50, seed=3)
    K = 30
    ref = OracleRIPTRM({"maxiter": K, "tolresid": 0, "manviofun": NonnegPCAProblem.manviofun}).run(
        NonnegPCAProblem(Z, x0, y0))
    x, y, sm, tr = detc.solve(Z, x0, y0, {"maxiter": K, "tolresid": 0}, trace_capacity=512)
    L = rb.trace_to_log(tr)
    first = first_discrete_mismatch(L, ref.log)
    assert ref.log["iteration"][min(first, len(ref.log["iteration"]) - 1)] >= 8
    a, b = _converged_costs(L), _converged_costs(ref.log)
    assert len(a) == len(b) == K
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-6
    assert abs(a[-1] - b[-1]) < REL_TOL * abs(b[-1])
    assert np.max(np.abs(x - ref.x)) < REL_TOL
    assert np.max(np.abs(y - ref.ineqLagmult)) < REL_TOL * max(1.0, np.max(np.abs(ref.ineqLagmult)))
    assert L["residual"][-1] < 1e-9 and ref.log["residual"][-1] < 1e-9


def test_c_oracle_hessvec_matches_per_constraint_operators():
    from oracle import riptrm_oracle as O
    rng = np.random.RandomState(1)
    Z, x0, _ = nonnegpca_generate_instance(50, seed=9)
    y = 0.5 + rng.rand(50)
    P = NonnegPCAProblem(Z, x0, y)
    v = P.manifold.projection(x0, rng.randn(50))
    s = O.slack(P, x0)
    ref = O.hess_lagrangian(P, x0, y, v) + O.G_apply(P, x0, (y * O.Gadj_apply(P, x0, v)) / s)
    out = detc.hessvec(Z, x0, y, 0.05, v)
    assert np.max(np.abs(out - ref)) < 1e-11 * np.max(np.abs(ref))


def test_c_oracle_is_deterministic_and_threads_agree():
    Z = np.stack([nonnegpca_generate_instance(50, seed=s)[0] for s in range(6)])
    X = np.stack([nonnegpca_generate_instance(50, seed=s)[1] for s in range(6)])
    Y = np.ones((6, 50))
    opt = {"maxiter": 6, "tolresid": 0}
    x1, y1, s1 = detc.solve_many(Z, X, Y, opt, threads=1)
    x2, y2, s2 = detc.solve_many(Z, X, Y, opt, threads=3)
    assert np.array_equal(x1, x2) and np.array_equal(y1, y2) and np.array_equal(s1, s2)


def test_merged_reduction_tcg_matches_reference_operation_order(datasets):
    """The kernel's merged-reduction tCG vs the same C oracle compiled with the tCG loop in the reference's
    operation order (-DFAITHFUL_TCG): same discrete trace in the well-conditioned window, same final iterate."""
    d = datasets["NonnegPCA/1"]
    opt = {"maxiter": 40, "tolresid": 0}
    xa, ya, sa, ta = detc.solve(d["Z"], d["initx_a"], d["initineqLagmult"], opt, trace_capacity=512)
    xb, yb, sb, tb = detc.solve(d["Z"], d["initx_a"], d["initineqLagmult"], opt, trace_capacity=512, faithful=True)
    La, Lb = rb.trace_to_log(ta), rb.trace_to_log(tb)
    first = first_discrete_mismatch(La, Lb, columns=DISCRETE_COLUMNS + ("tcg_iters",))
    assert Lb["iteration"][min(first, len(Lb["iteration"]) - 1)] > 8
    assert np.max(np.abs(xa - xb)) < 1e-12 and abs(sa[0] - sb[0]) < 1e-12 * abs(sb[0])
    g = load_golden("nonnegpca_1_a_K40")
    assert np.max(np.abs(xb - np.array(g["x"]))) < REL_TOL


def test_model_decrease_from_the_tcg_product_moves_no_decision(datasets, monkeypatch):
    """inner_step evaluates pred = -<c,dx> - <dx,Hw dx>/2 (RIPTRM.py:659-660) with the Hw[eta] the tCG accumulated beside eta;
    RIPTRM_RECOMPUTE_HDX=1 forms the fresh product the reference writes.  Hw is linear, so the two differ by rounding only:
    on the reference's dataset every discrete column of the 343-row trace, every radius and every tCG count is the same, the
    objective agrees to 1e-13 and `ared_pred` to 1e-6 wherever it is not a ratio of two rounding-level numbers."""
    d = datasets["NonnegPCA/1"]
    runs = []
    for fresh in (False, True):
        if fresh:
            monkeypatch.setenv("RIPTRM_RECOMPUTE_HDX", "1")
        else:
            monkeypatch.delenv("RIPTRM_RECOMPUTE_HDX", raising=False)
        x, y, sm, tr = detc.solve(d["Z"], d["initx_a"], d["initineqLagmult"], {"maxiter": 40, "tolresid": 0},
                                  trace_capacity=512)
        runs.append((x, y, sm, rb.trace_to_log(tr)))
    (xa, ya, sa, La), (xb, yb, sb, Lb) = runs
    assert sa[13] == 0 and sb[13] > 0                                # fresh products formed: none / one per rho test
    assert len(La["iteration"]) == len(Lb["iteration"]) == 343   # row 0 + 342 trust-region iterations
    assert first_discrete_mismatch(La, Lb, columns=DISCRETE_COLUMNS + ("tcg_iters",)) == len(La["iteration"])
    assert max_rel_diff(La, Lb, "TR_radius") == 0.0
    assert max_rel_diff(La, Lb, "cost") < 1e-13
    assert np.max(np.abs(xa - xb)) < 1e-14 and sa[12] == sb[12]
    ra = np.array([v if v is not None else np.nan for v in La["ared/pred"]], dtype=float)
    rbv = np.array([v if v is not None else np.nan for v in Lb["ared/pred"]], dtype=float)
    ok = np.isfinite(ra) & np.isfinite(rbv) & (np.array(La["iteration"]) <= 30)
    assert np.nanmax(np.abs(ra[ok] - rbv[ok]) / np.maximum(1.0, np.abs(rbv[ok]))) < 1e-6
