"""GPU parity of the Rosenbrock / Grassmann(5,3) family (BASELINE config 2) against the NumPy oracle and the
unmodified reference's golden run (tests/golden/rosenbrock_K6.json).  alpha = 1e7 makes the problem chaotic in
rounding (the NumPy oracle itself leaves the reference's trace at row 45), so whole-run comparisons use a window."""
import numpy as np
import pytest

from conftest import load_golden
from helpers import DISCRETE_COLUMNS, first_discrete_mismatch, max_rel_diff, rosenbrock_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import riptrm_b200
    return riptrm_b200


def _structure(rb, x0=None, y0=None):
    P = rosenbrock_problem()
    return rb.RosenbrockStructure(n=5, k=3, alpha=1e7, x0=P.initialpoint if x0 is None else x0,
                                  y0=P.initialineqLagmult if y0 is None else y0), P


def _interior_point(rng):
    """A feasible point of Grassmann(5,3) with all entries > -0.01 and a random tangent vector."""
    from oracle.manifolds import Grassmann
    man = Grassmann(5, 3)
    X = np.eye(5)[:, :3].copy()
    X[3:, :] += 0.03 * rng.rand(2, 3)
    X = man.retraction(X, np.zeros((5, 3)))
    assert (X > -0.01).all()
    return man, X


def test_iteration0_known_answer(rb):
    """src/Rosenbrock/analyzer.ipynb cell 5 row 0: cost 5.000001e+07, residual 2.000000e+07, compl 1.749714."""
    st, _ = _structure(rb)
    out = rb.RIPTRM({"TRS_solver": "tCG", "second_order_stationarity": False, "maxiter": 0, "maxtime": 1e9}).run_batch(
        [None], structures=[st])[0]
    assert abs(out.log["cost"][0] / 5.000001e7 - 1) < 1e-6
    assert abs(out.log["residual"][0] / 2.0e7 - 1) < 1e-6
    assert abs(out.log["complviolation"][0] - 1.749714) < 1e-6
    assert out.log["manviolation"][0] == 0.0


def test_hessvec_and_tcg_hooks_match_oracle(rb):
    from oracle import riptrm_oracle as O
    from oracle.problems import RosenbrockProblem
    rng = np.random.RandomState(5)
    sts, probs, V = [], [], []
    for _ in range(4):
        man, X = _interior_point(rng)
        y = 0.5 + rng.rand(15)
        P = RosenbrockProblem(5, 3, 1e7)
        P.initialpoint, P.initialineqLagmult = X, y
        probs.append(P)
        sts.append(rb.RosenbrockStructure(n=5, k=3, alpha=1e7, x0=X, y0=y))
        V.append(man.projection(X, rng.randn(5, 3)))
    bs = rb.BatchSolver(sts)
    mu = 0.1
    hv = bs.hessvec(bs.x0, bs.y0, mu, np.array([v.reshape(-1) for v in V]))
    for i, P in enumerate(probs):
        x, y = P.initialpoint, P.initialineqLagmult
        s = O.slack(P, x)
        ref = O.hess_lagrangian(P, x, y, V[i]) + O.G_apply(P, x, (y * O.Gadj_apply(P, x, V[i])) / s)
        assert np.max(np.abs(hv[i].reshape(5, 3) - ref)) < 1e-10 * np.max(np.abs(ref))
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=1)
    bs.set_options(opt)
    for Delta in (1e-3, 0.2):
        eta, info = bs.tcg(bs.x0, bs.y0, mu, Delta)
        for i, P in enumerate(probs):
            x, y = P.initialpoint, P.initialineqLagmult
            s = O.slack(P, x)
            Hw = lambda _x, dx: O.hess_lagrangian(P, x, y, dx) + O.G_apply(P, x, (y * O.Gadj_apply(P, x, dx)) / s)
            c = P.riemannian_gradient(x) - O.G_apply(P, x, mu / s)
            e_ref, _, j, stop = O.steihaug_tcg(P.manifold, Hw, x, c, Delta, 1, 0.1, 1, P.manifold.dim, P.preconditioner)
            assert int(info[i, 0]) == j + 1 and O.TCG_STOPS[int(info[i, 1])] == stop
            assert np.max(np.abs(eta[i].reshape(5, 3) - e_ref)) < 1e-7 * max(1e-6, np.max(np.abs(e_ref)))
            assert np.linalg.norm(eta[i]) <= Delta * (1 + 1e-12)
    bs.close()


def test_trace_matches_reference_and_oracle_window(rb):
    """First outer iteration, 60 trust-region iterations: the discrete trace (status, tCG stop reason and count, radius
    update) equals the reference's golden run for the first 40 rows, the objective agrees to 1e-9 there."""
    from oracle.problems import RosenbrockProblem
    from oracle.riptrm_oracle import OracleRIPTRM
    g = load_golden("rosenbrock_K6")
    G = dict(g["log"], tcg_iters=[None] + g["tcg_iters"])
    st, P = _structure(rb)
    out = rb.RIPTRM({"TRS_solver": "tCG", "second_order_stationarity": False, "maxiter": 1, "inner_maxiter": 60,
                     "tolresid": 0, "maxtime": 1e9}).run_batch([None], structures=[st])[0]
    first = first_discrete_mismatch(out.log, G, columns=DISCRETE_COLUMNS + ("tcg_iters",))
    assert first >= 40, first
    assert max_rel_diff(out.log, G, "cost", rows=first) < 1e-9
    assert max_rel_diff(out.log, G, "TR_radius", rows=first) < 1e-9
    ref = OracleRIPTRM({"maxiter": 1, "inner_maxiter": 60, "tolresid": 0, "manviofun": RosenbrockProblem.manviofun}).run(P)
    first_o = first_discrete_mismatch(out.log, ref.log, columns=DISCRETE_COLUMNS + ("tcg_iters",))
    assert first_o >= 40, first_o
    assert max_rel_diff(out.log, ref.log, "cost", rows=first_o) < 1e-9
    assert max_rel_diff(out.log, ref.log, "normdx", rows=20) < 1e-6


def test_initial_points_abc_are_identical_runs(rb):
    """BASELINE config 2 'initial points a, b, c': the coordinator ignores the name (src/Rosenbrock/coordinator.py:78-84),
    so the three runs must be bit-identical -- a determinism check of the batched kernel."""
    st, _ = _structure(rb)
    outs = rb.RIPTRM({"TRS_solver": "tCG", "second_order_stationarity": False, "maxiter": 3, "tolresid": 0,
                      "maxtime": 1e9}).run_batch([None] * 3, structures=[st, st, st])
    for o in outs[1:]:
        assert np.array_equal(o.x, outs[0].x) and np.array_equal(o.ineqLagmult, outs[0].ineqLagmult)
        assert o.log["cost"] == outs[0].log["cost"] and o.log["dxtype"] == outs[0].log["dxtype"]
    assert outs[0].log["cost"][-1] < 4.1e7   # the golden run reaches 4.0000009514e7 by outer iteration 14
