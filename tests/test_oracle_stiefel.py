"""NonnegPCA on Stiefel(n, p) (extrapolated workload: BASELINE config 4 as written): the matrix-form
operators used as the oracle for the CUDA STIEFEL family against the per-constraint restatement of the reference's
operators (RIPTRM.py:457-571) on the pymanopt-formula Stiefel manifold, at a size the per-constraint path can afford."""
import numpy as np
import pytest

from oracle import riptrm_oracle as ro
from oracle.problems import NonnegPCAStiefelProblem, nonnegpca_generate_Z
from helpers import stiefel_start


def _pair(n=24, p=3, seed=5, eps=0.01):
    Z, _ = nonnegpca_generate_Z(n, seed=seed)
    X0 = stiefel_start(n, p, seed)
    rs = np.random.RandomState(seed + 1)
    y = 0.5 + rs.rand(n * p)
    return (NonnegPCAStiefelProblem(Z, X0, y, eps=eps), NonnegPCAStiefelProblem(Z, X0, y, eps=eps, closed_form=True), rs)


def test_start_point_is_feasible_and_orthonormal():
    X0 = stiefel_start(40, 4, 1)
    assert np.allclose(X0.T @ X0, np.eye(4), atol=1e-14) and X0.min() >= 0.0


@pytest.mark.parametrize("emb", [False, True])
def test_closed_form_operators_match_per_constraint_operators(emb):
    gen, clo, rs = _pair()
    X, y, man = gen.initialpoint, gen.initialineqLagmult, gen.manifold
    V = man.projection(X, rs.randn(*X.shape))
    w = rs.rand(y.size)
    rel = lambda a, b: np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1e-300, np.max(np.abs(np.asarray(b))))
    assert rel(ro.slack(clo, X), ro.slack(gen, X)) == 0.0
    assert rel(ro.grad_lagrangian(clo, X, y), ro.grad_lagrangian(gen, X, y)) < 1e-13
    assert rel(ro.hess_lagrangian(clo, X, y, V), ro.hess_lagrangian(gen, X, y, V)) < 1e-13
    assert rel(ro.G_apply(clo, X, w), ro.G_apply(gen, X, w)) < 1e-13
    assert rel(ro.Gadj_apply(clo, X, V, emb), ro.Gadj_apply(gen, X, V, emb)) < 1e-13
    a, b = ro.kkt_residual(clo, X, y, NonnegPCAStiefelProblem.manviofun), ro.kkt_residual(gen, X, y, NonnegPCAStiefelProblem.manviofun)
    assert rel(a, b) < 1e-13


def test_closed_form_run_matches_per_constraint_run():
    """Whole RIPTRM runs (6 outer iterations, some 90 trust-region iterations) through both operator paths: the discrete
    trace is identical well into the run (the two paths differ by rounding only, and long tCG runs amplify that --
    SURVEY.md App. C), and the outer iterations arrive at the same cost and point."""
    gen, clo, _ = _pair(n=16, p=2, seed=9)
    outs = []
    for pb in (gen, clo):
        o = ro.OracleRIPTRM({"maxiter": 6, "tolresid": 0, "manviofun": NonnegPCAStiefelProblem.manviofun})
        outs.append(o.run(pb))
    a, b = outs
    same = [i for i, (u, v, w, z) in enumerate(zip(a.log["inner_status"], b.log["inner_status"], a.log["dxtype"], b.log["dxtype"]))
            if (u, w) != (v, z)]
    assert not same or same[0] > 40
    conv = lambda L: np.array([c for c, st in zip(L["cost"], L["inner_status"]) if st == "converged"])
    assert len(conv(a.log)) == len(conv(b.log)) == 6
    # early outer iterations end at loose tolerances (mu = 0.1, 0.05, ...) wherever the inner loop happens to stop
    assert np.max(np.abs(conv(a.log) - conv(b.log))) < 1e-3 and abs(conv(a.log)[-1] - conv(b.log)[-1]) < 1e-8 * abs(conv(b.log)[-1])
    assert np.max(np.abs(a.x - b.x)) < 1e-6
    assert np.isnan(a.log["distance"][-1])
