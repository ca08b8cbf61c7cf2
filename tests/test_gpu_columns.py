"""GPU parity of the large-n COLUMNS family (BASELINE config 4 at sizes the oracle finishes in seconds):
the cooperative stream-K Hessian-vector kernel and the lock-step tCG against the per-column Sphere oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import riptrm_b200
    return riptrm_b200


def _instance(n, p, seed):
    from oracle.problems import nonnegpca_generate_Z
    Z, rs = nonnegpca_generate_Z(n, seed=seed)
    X = rs.rand(n, p)
    X = np.abs(X / np.linalg.norm(X, axis=0, keepdims=True))
    Y = 0.5 + rs.rand(n, p)
    return Z, X, Y, rs


def _closed_form_hw(Z, x, y, mu, v):
    """Hw[v] of SURVEY.md App. A.1 in NumPy (validated against the per-constraint oracle in test_oracle_*)."""
    S = Z + Z.T
    P = lambda u: u - (x @ u) * x
    return P(-S @ v) + (x @ S @ x + y @ x) * v + P((y / x) * (v - x * (x @ v)))


@pytest.mark.parametrize("n,p", [(96, 1), (500, 4), (1000, 10), (777, 3), (640, 16), (1000, 12)])
def test_hessvec_matches_closed_form_and_oracle(rb, n, p):
    from oracle import riptrm_oracle as O
    from oracle.problems import NonnegPCAProblem
    Z, X, Y, rs = _instance(n, p, seed=n + p)
    V = rs.randn(n, p)
    V -= X * np.sum(X * V, axis=0, keepdims=True)
    cs = rb.ColumnsSolver(Z, p)
    out = cs.hessvec(X, Y, 0.05, V)
    assert cs.matvec_passes == 2
    for c in range(p):
        ref = _closed_form_hw(Z, X[:, c], Y[:, c], 0.05, V[:, c])
        assert np.max(np.abs(out[:, c] - ref)) < 1e-10 * max(1.0, np.max(np.abs(ref)))
    if n <= 100:  # the per-constraint oracle costs O(n^2) Python calls per product
        Pb = NonnegPCAProblem(Z, X[:, 0], Y[:, 0])
        x, y = X[:, 0], Y[:, 0]
        s = O.slack(Pb, x)
        ref = O.hess_lagrangian(Pb, x, y, V[:, 0]) + O.G_apply(Pb, x, (y * O.Gadj_apply(Pb, x, V[:, 0])) / s)
        assert np.max(np.abs(out[:, 0] - ref)) < 1e-10 * max(1.0, np.max(np.abs(ref)))
    cs.close()


@pytest.mark.parametrize("n,p,Delta", [(300, 4, 0.3), (300, 4, 5.0), (1000, 10, 0.05), (641, 2, 1.0), (520, 16, 0.5)])
def test_lockstep_tcg_matches_oracle_tcg(rb, n, p, Delta):
    """Each column's tCG (iteration count, stop reason, eta) equals the oracle's Steihaug-Toint tCG
    (RIPTRM.py:41-216) run on that column with the closed-form operator."""
    from oracle import riptrm_oracle as O
    from oracle.manifolds import Sphere
    Z, X, Y, rs = _instance(n, p, seed=3 * n + p)
    mu = 0.1
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=1)
    cs = rb.ColumnsSolver(Z, p, option=opt)
    eta, info = cs.tcg(X, Y, mu, Delta)
    S = Z + Z.T
    man = Sphere(n)
    for c in range(p):
        x, y = X[:, c], Y[:, c]
        Pj = lambda u: u - (x @ u) * x
        grad = Pj(-S @ x) - Pj(mu / x)
        e_ref, _, j, stop = O.steihaug_tcg(man, lambda _x, v: _closed_form_hw(Z, x, y, mu, v), x, grad, Delta, 1, 0.1, 1,
                                           man.dim, lambda _x, v: v)
        assert int(info[c, 0]) == j + 1, (c, info[c], j + 1, stop)
        assert O.TCG_STOPS[int(info[c, 1])] == stop
        assert np.max(np.abs(eta[:, c] - e_ref)) < 1e-8 * max(1e-3, np.max(np.abs(e_ref)))
        assert abs(info[c, 2] - np.linalg.norm(e_ref)) < 1e-8 * max(1e-3, np.linalg.norm(e_ref))
        assert abs(x @ eta[:, c]) < 1e-10
    cs.close()


def test_streaming_pass_is_linear_and_symmetric(rb):
    """Size-independent properties at a larger n: Hw is linear in V and self-adjoint on the tangent space."""
    n, p = 4096, 10
    Z, X, Y, rs = _instance(n, p, seed=11)
    cs = rb.ColumnsSolver(Z, p)
    proj = lambda V: V - X * np.sum(X * V, axis=0, keepdims=True)
    U, V = proj(rs.randn(n, p)), proj(rs.randn(n, p))
    HU, HV = cs.hessvec(X, Y, 0.01, U), cs.hessvec(X, Y, 0.01, V)
    HUV = cs.hessvec(X, Y, 0.01, U + 2.0 * V)
    scale = np.max(np.abs(HUV))
    assert np.max(np.abs(HUV - (HU + 2.0 * HV))) < 1e-11 * scale
    uhv, vhu = np.sum(U * HV, axis=0), np.sum(V * HU, axis=0)
    assert np.max(np.abs(uhv - vhu)) < 1e-9 * np.max(np.abs(uhv))
    # determinism: same bits run to run
    assert np.array_equal(cs.hessvec(X, Y, 0.01, U), HU)
    cs.close()


def test_whole_solve_matches_the_c_oracle_per_column(rb):
    """riptrm_solve on the COLUMNS family: every column is an independent Sphere RIPTRM run.  n = 100, p = 4 against
    the C oracle (same merged-reduction tCG, different summation order) run column by column: identical outer / inner / tCG iteration counts in
    the well-conditioned window (8 outer iterations), final iterate / objective / multipliers to 1e-8 after 30."""
    from oracle.c import binding as detc
    n, p = 100, 4
    Z, X0, Y0, rs = _instance(n, p, seed=77)
    Y0 = np.ones((n, p))
    cs = rb.ColumnsSolver(Z, p)
    for K, exact_counts in ((8, True), (30, False)):
        opt = rb.options.default_option()
        opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=K, tolresid=0, maxtime=1e9, inner_maxiter=1000)
        X, Y, sm, tr = cs.solve(X0, Y0, opt, per_outer_trace=True)
        SM = rb._lib.SM
        for c in range(p):
            xo, yo, so, _ = detc.solve(Z, X0[:, c].copy(), Y0[:, c].copy(), {"maxiter": K, "tolresid": 0, "inner_maxiter": 1000})
            assert sm[c, SM["outer_iters"]] == so[10] == K and sm[c, SM["stop_reason"]] == 2
            if exact_counts:
                assert sm[c, SM["inner_iters"]] == so[11], (c, sm[c, SM["inner_iters"]], so[11])
                # tCG counts hinge on the ulp-level test norm_r <= target (RIPTRM.py:183): a different summation
                # order may stop one iteration apart
                assert abs(sm[c, SM["tcg_iters"]] - so[12]) <= 2, (c, sm[c, SM["tcg_iters"]], so[12])
                # the Hw[dx] of RIPTRM.py:659 is the product the tCG accumulated, in the kernels and in the C oracle
                assert so[13] == 0 and sm[c, SM["aux_hessvecs"]] == 0
            assert abs(sm[c, SM["cost"]] - so[0]) < (1e-6 if exact_counts else 1e-8) * abs(so[0])
            if not exact_counts:
                assert np.max(np.abs(X[:, c] - xo)) < 1e-8
                assert np.max(np.abs(Y[:, c] - yo)) < 1e-8 * max(1.0, np.max(np.abs(yo)))
                assert sm[c, SM["residual"]] < 1e-9 and so[1] < 1e-9
            assert abs(np.linalg.norm(X[:, c]) - 1) < 1e-14 and (X[:, c] > 0).all()
            # per-outer trace rows: iteration 0..K, cost decreasing to the final value
            rows = int(sm[c, SM["trace_rows"]])
            assert rows == K + 1
            assert list(tr[c, :rows, rb._lib.TR["iteration"]]) == list(range(K + 1))
            assert tr[c, rows - 1, rb._lib.TR["cost"]] == sm[c, SM["cost"]]
    cs.close()


def test_whole_solve_columns_finish_independently(rb):
    """Columns stop at their own time (tolresid reached at different outer iterations); finished columns are frozen."""
    n, p = 300, 3
    Z, X0, Y0, rs = _instance(n, p, seed=5)
    Y0 = np.ones((n, p))
    cs = rb.ColumnsSolver(Z, p)
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=40, tolresid=1e-6, maxtime=1e9, inner_maxiter=1000)
    X, Y, sm, _ = cs.solve(X0, Y0, opt)
    SM = rb._lib.SM
    assert (sm[:, SM["stop_reason"]] == 3).all() and (sm[:, SM["residual"]] <= 1e-6).all()
    assert (sm[:, SM["outer_iters"]] < 40).all()
    cs.close()


def test_run_columns_returns_reference_style_outputs(rb):
    """RIPTRM(option).run_columns: one Output per column with the reference's per-outer-iteration log layout."""
    n, p = 128, 3
    Z, X0, Y0, rs = _instance(n, p, seed=21)
    solver = rb.RIPTRM({"TRS_solver": "tCG", "second_order_stationarity": False, "maxiter": 12, "tolresid": 0, "maxtime": 1e9,
                        "save_inner_iteration": False})
    outs = solver.run_columns(Z, X0, np.ones((n, p)))
    assert len(outs) == p
    for c, o in enumerate(outs):
        assert o.name == "RIPTRM_tCG" and o.x.shape == (n,) and o.ineqLagmult.shape == (n,)
        assert o.log["iteration"] == list(range(13))
        assert list(o.log)[:11] == ["iteration", "time", "cost", "distance", "residual", "gradnorm", "complviolation",
                                    "dualviolation", "manviolation", "maxviolation", "meanviolation"]
        assert all(b <= a + 1e-12 for a, b in zip(o.log["cost"][1:], o.log["cost"][2:]))   # objective decreases
        assert o.option["stoppingcriterion"].startswith("Max iteration count reached; maxiter=12")
        assert abs(np.linalg.norm(o.x) - 1) < 1e-14


def test_per_inner_iteration_log_matches_the_c_oracle(rb):
    """run_columns with the reference's default save_inner_iteration=True: one log row per trust-region iteration and
    column.  Against the C oracle's trace of the same column (n = 100): identical discrete columns (status, tCG stop reason,
    radius update, clipping flag) and radii in the well-conditioned window, float columns to 1e-8 there."""
    from helpers import DISCRETE_COLUMNS, first_discrete_mismatch, max_rel_diff
    from oracle.c import binding as detc
    n, p = 100, 3
    Z, X0, Y0, rs = _instance(n, p, seed=31)
    Y0 = np.ones((n, p))
    K = 10
    solver = rb.RIPTRM({"TRS_solver": "tCG", "second_order_stationarity": False, "maxiter": K, "tolresid": 0, "maxtime": 1e9})
    outs = solver.run_columns(Z, X0, Y0)
    for c, o in enumerate(outs):
        xo, yo, so, tr = detc.solve(Z, X0[:, c].copy(), Y0[:, c].copy(), {"maxiter": K, "tolresid": 0}, trace_capacity=400)
        ref = rb.trace_to_log(tr)
        assert o.log["iteration"][0] == 0 and o.log["inner_status"][0] is None
        first = first_discrete_mismatch(o.log, ref)
        outer = ref["iteration"][min(first, len(ref["iteration"]) - 1)]
        assert outer >= 7 or first == len(ref["iteration"]), (c, first, outer)
        m = min(first, 25)
        assert max_rel_diff(o.log, ref, "TR_radius", rows=m) < 1e-12
        for col in ("cost", "mu", "normdx", "maxabsLagmult", "minxfeasi", "compl"):
            assert max_rel_diff(o.log, ref, col, rows=m, floor=1e-12) < 1e-6, (c, col)
        assert max_rel_diff(o.log, ref, "residual", rows=m, floor=1e-10) < 1e-5
        assert abs(o.log["cost"][-1] - ref["cost"][-1]) < 1e-7 * abs(ref["cost"][-1])


@pytest.mark.parametrize("family", ["columns", "stiefel"])
def test_device_side_loop_and_host_sequenced_loop_agree_bit_for_bit(rb, family, monkeypatch):
    """The whole solve runs as one graph launch with a conditional WHILE node (device-side loop); RIPTRM_COLUMNS_HOST_LOOP=1
    sequences the same launches from the host.  Same launches, same order: identical iterates, multipliers, summaries, logs."""
    n, p = 100, 4
    Z, X0, Y0, rs = _instance(n, p, seed=5)
    Y0 = np.ones((n, p))
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=6, tolresid=0, maxtime=1e9, inner_maxiter=50)
    res = []
    for host_loop in (False, True):
        if host_loop:
            monkeypatch.setenv("RIPTRM_COLUMNS_HOST_LOOP", "1")
        else:
            monkeypatch.delenv("RIPTRM_COLUMNS_HOST_LOOP", raising=False)
        if family == "columns":
            s = rb.ColumnsSolver(Z, p)
            x0 = X0
        else:
            from helpers import stiefel_start
            s = rb.StiefelSolver(Z, p, eps=0.01)
            x0 = stiefel_start(n, p, 5)
        res.append(s.solve(x0, Y0, opt, per_inner_trace=True, trace_capacity=600))
        s.close()
    (Xa, Ya, sa, ta), (Xb, Yb, sb, tb) = res
    TIME = rb._lib.TR["time"]
    keep = [i for i in range(ta.shape[2]) if i != TIME]
    assert np.array_equal(Xa, Xb) and np.array_equal(Ya, Yb) and np.array_equal(sa, sb)
    assert np.array_equal(np.nan_to_num(ta[:, :, keep], nan=-7.0), np.nan_to_num(tb[:, :, keep], nan=-7.0))
    assert sa[0, rb._lib.SM["outer_iters"]] >= 1
