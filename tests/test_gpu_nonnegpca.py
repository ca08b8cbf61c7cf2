"""GPU parity: CUDA NonnegPCA/Sphere path (through the C ABI) vs the NumPy oracle and the
reference's own golden run (tests/golden/nonnegpca_1_a_K40.json)."""
import numpy as np
import pytest

from conftest import load_golden
from helpers import (DISCRETE_COLUMNS, first_discrete_mismatch, max_rel_diff, nonnegpca_problem)

pytestmark = pytest.mark.gpu

REL_TOL = 1e-8  # north_star: objective, KKT residual and iterates to a relative 1e-8 in fp64


@pytest.fixture(scope="module")
def rb():
    import riptrm_b200
    return riptrm_b200


def _solve(rb, st, **opt):
    option = {"TRS_solver": "tCG", "second_order_stationarity": False, "tolresid": 0, "maxtime": 1e9}
    option.update(opt)
    solver = rb.RIPTRM(option)
    return solver.run_batch([None], structures=[st])[0], solver


def test_reference_dataset_trace_matches_golden(rb, datasets):
    """NonnegPCA instance 1 / initial point a (BASELINE config 1), 40 outer iterations: every discrete
    log column (inner status, tCG stop reason, radius update, ...) and the tCG iteration counts are
    identical to the reference's run inside the well-conditioned window, and the objective / iterates
    agree to 1e-8 through the whole run."""
    g = load_golden("nonnegpca_1_a_K40")
    d = datasets["NonnegPCA/1"]
    st = rb.NonnegPCAStructure(Z=d["Z"], x0=d["initx_a"], y0=d["initineqLagmult"])
    out, solver = _solve(rb, st, maxiter=40)
    L, G = out.log, g["log"]
    G = dict(G, tcg_iters=[None] + g["tcg_iters"])
    nrows = len(G["iteration"])
    outer_of = lambda row: G["iteration"][min(row, nrows - 1)] if row < nrows else 10 ** 9
    # discrete columns (inner status, tCG stop reason, radius update, clipping flag): SURVEY App. C
    # found rounding-order changes leave outer iterations 1-16 identical on this instance
    first = first_discrete_mismatch(L, G)
    assert outer_of(first) > 16, f"discrete trace diverges at row {first} (outer {outer_of(first)})"
    # tCG iteration counts hinge on `norm_r <= target` (RIPTRM.py:183), an ulp-level test
    first_tcg = first_discrete_mismatch(L, G, columns=DISCRETE_COLUMNS + ("tcg_iters",))
    assert outer_of(first_tcg) > 8, f"tCG counts diverge at row {first_tcg} (outer {outer_of(first_tcg)})"
    # until the first differing tCG count the trust-region radii are exact; inner iterates differ transiently
    # (long tCG runs amplify rounding) and re-converge at every outer iteration
    assert max_rel_diff(L, G, "TR_radius", rows=first_tcg) == 0.0
    assert max_rel_diff(L, G, "mu", rows=first) == 0.0
    assert max_rel_diff(L, G, "cost", rows=first) < 1e-6
    # per-outer-iteration cost (rows where the inner loop converged) through the whole run
    conv = lambda log: [c for c, s in zip(log["cost"], log["inner_status"]) if s == "converged"]
    a, b = np.array(conv(L)), np.array(conv(G))
    assert len(a) == len(b) == 40
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-6
    assert np.max(np.abs(a[-10:] - b[-10:]) / np.abs(b[-10:])) < REL_TOL
    assert np.max(np.abs(out.x - np.array(g["x"]))) < REL_TOL
    assert abs(L["cost"][-1] - G["cost"][-1]) < REL_TOL * abs(G["cost"][-1])
    assert L["residual"][-1] < 1e-10 and G["residual"][-1] < 1e-10


def test_iteration0_known_answer(rb, datasets):
    """Notebook known answer (src/NonnegPCA/analyzer.ipynb cell 5): iteration-0 residual 4.986888e+00."""
    d = datasets["NonnegPCA/1"]
    st = rb.NonnegPCAStructure(Z=d["Z"], x0=d["initx_a"], y0=d["initineqLagmult"])
    out, _ = _solve(rb, st, maxiter=0)
    assert abs(out.log["residual"][0] - 4.986888432851818) < 1e-12
    assert abs(out.log["cost"][0] - (-0.5093080157946566)) < 1e-14
    assert out.option["stoppingcriterion"].startswith("Max iteration count reached; maxiter=0")


def test_hessvec_and_tcg_hooks_match_oracle(rb, datasets):
    """riptrm_hessvec / riptrm_tcg against the oracle's per-constraint operators (RIPTRM.py:729, :41-216)."""
    from oracle import riptrm_oracle as O
    from oracle.problems import nonnegpca_generate_instance, NonnegPCAProblem
    rng = np.random.RandomState(7)
    sts, probs = [], []
    for seed in range(6):
        Z, x0, y0 = nonnegpca_generate_instance(50, seed=seed)
        y0 = 0.5 + rng.rand(50)
        sts.append(rb.NonnegPCAStructure(Z=Z, x0=x0, y0=y0))
        probs.append(NonnegPCAProblem(Z, x0, y0))
    bs = rb.BatchSolver(sts)
    mu = 0.05
    V = []
    for P in probs:
        v = rng.randn(50)
        V.append(P.manifold.projection(P.initialpoint, v))
    V = np.array(V)
    hv = bs.hessvec(bs.x0, bs.y0, mu, V)
    for i, P in enumerate(probs):
        x, y = P.initialpoint, P.initialineqLagmult
        s = O.slack(P, x)
        ref = O.hess_lagrangian(P, x, y, V[i]) + O.G_apply(P, x, (y * O.Gadj_apply(P, x, V[i])) / s)
        assert np.max(np.abs(hv[i] - ref)) < 1e-11 * max(1.0, np.max(np.abs(ref)))
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=1)
    bs.set_options(opt)
    for Delta in (0.05, 0.4, 3.0):
        eta, info = bs.tcg(bs.x0, bs.y0, mu, Delta)
        for i, P in enumerate(probs):
            x, y = P.initialpoint, P.initialineqLagmult
            s = O.slack(P, x)
            Hw = lambda _x, dx: O.hess_lagrangian(P, x, y, dx) + O.G_apply(P, x, (y * O.Gadj_apply(P, x, dx)) / s)
            c = P.riemannian_gradient(x) - O.G_apply(P, x, mu / s)
            e_ref, _, j, stop = O.steihaug_tcg(P.manifold, Hw, x, c, Delta, 1, 0.1, 1, P.manifold.dim, P.preconditioner)
            assert int(info[i, 0]) == j + 1
            assert O.TCG_STOPS[int(info[i, 1])] == stop
            assert np.max(np.abs(eta[i] - e_ref)) < 1e-9 * max(1e-3, np.max(np.abs(e_ref)))
            assert np.linalg.norm(eta[i]) <= Delta * (1 + 1e-12)
            assert abs(x @ eta[i]) < 1e-12  # tangency
    bs.close()


def test_batch_matches_oracle_on_generated_instances(rb):
    """Config-5 style instances (reference generator law, seed = instance id), 30 outer iterations:
    identical per-outer inner-iteration counts inside the window, objective / iterate parity at the end."""
    from oracle.problems import nonnegpca_generate_instance, NonnegPCAProblem
    from oracle.riptrm_oracle import OracleRIPTRM
    sts, outs_ref = [], []
    for seed in range(2):
        Z, x0, y0 = nonnegpca_generate_instance(50, seed=100 + seed)
        sts.append(rb.NonnegPCAStructure(Z=Z, x0=x0, y0=y0))
        outs_ref.append(OracleRIPTRM({"maxiter": 30, "tolresid": 0, "manviofun": NonnegPCAProblem.manviofun}).run(
            NonnegPCAProblem(Z, x0, y0)))
    solver = rb.RIPTRM({"TRS_solver": "tCG", "second_order_stationarity": False, "tolresid": 0, "maxtime": 1e9,
                        "maxiter": 30})
    outs = solver.run_batch([None] * 2, structures=sts)
    for o, r in zip(outs, outs_ref):
        first = first_discrete_mismatch(o.log, r.log)
        outer = r.log["iteration"][min(first, len(r.log["iteration"]) - 1)]
        assert outer >= 8 or first == len(r.log["iteration"])
        assert abs(o.log["cost"][-1] - r.log["cost"][-1]) < REL_TOL * abs(r.log["cost"][-1])
        assert np.max(np.abs(o.x - r.x)) < REL_TOL
        assert o.log["residual"][-1] < 1e-9


def test_gpu_trace_is_bit_identical_to_the_c_oracle(rb, datasets):
    """Parity tier T1: the CUDA kernel and the deterministic C oracle implement the same
    arithmetic specification (reduction tree, fma policy), so the WHOLE 40-outer-iteration trace -- every
    discrete column, every tCG iteration count and every floating-point column -- is equal bit for bit
    (`distance` goes through acos() of two different math libraries: 1e-12; `time` is wall clock)."""
    from oracle.c import binding as detc
    from riptrm_b200 import _lib
    d = datasets["NonnegPCA/1"]
    st = rb.NonnegPCAStructure(Z=d["Z"], x0=d["initx_a"], y0=d["initineqLagmult"])
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, tolresid=0, maxtime=1e9, maxiter=40)
    bs = rb.BatchSolver([st])
    bs.set_options(opt, 1, 512)
    x, y, sm, tr = bs.solve()
    bs.close()
    xo, yo, smo, tro = detc.solve(d["Z"], d["initx_a"], d["initineqLagmult"], {"maxiter": 40, "tolresid": 0},
                                  trace_capacity=512)
    rows = int(sm[0, _lib.SM["trace_rows"]])
    assert rows == len(tro) == int(smo[15])
    T = _lib.TR
    exact = [i for name, i in T.items() if name not in ("time", "distance", "mineigvalHw")]   # the C oracle writes 25 fields
    assert np.isnan(tr[0, :rows, T["mineigvalHw"]]).all()                                      # (no eigenvalue test under tCG)
    a, b = tr[0, :rows][:, exact], tro[:, exact]
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"first differing (row, field): {np.argwhere(~same)[:5]}"
    assert np.allclose(tr[0, :rows, T["distance"]], tro[:, T["distance"]], rtol=0, atol=1e-12)
    assert np.array_equal(x[0], xo) and np.array_equal(y[0], yo)
    assert np.array_equal(sm[0, :15], smo[:15])


@pytest.mark.parametrize("split", [-1, 5])
def test_gpu_batch_is_bit_identical_to_the_c_oracle_on_generated_pairs(rb, split):
    """64 pairs of the bench workload (config 5), full protocol: x, y and every summary field equal bit for bit,
    in one launch (split -1) and with the two-launch longest-first schedule forced (pause after 5 outer iterations)."""
    from oracle.c import binding as detc
    B = 64
    Z, x0, y0 = rb.datagen.nonnegpca_batch(1000, B, 50)
    opt = rb.options.default_option()
    proto = dict(TRS_solver="tCG", second_order_stationarity=False, maxiter=30, inner_maxiter=1000, tolresid=0,
                 maxtime=1e9)
    opt.update(proto)
    opt["schedule_split"] = split
    bs = rb.BatchSolver.nonnegpca_from_arrays(Z, x0, y0)
    bs.set_options(opt, 0, 0)
    x, y, sm, _ = bs.solve()
    bs.close()
    xo, yo, smo = detc.solve_many(Z, x0, y0, {"maxiter": 30, "inner_maxiter": 1000, "tolresid": 0}, threads=4)
    assert np.array_equal(x, xo) and np.array_equal(y, yo)
    assert np.array_equal(sm[:, :15], smo[:, :15])
    assert (sm[:, 1] < 1e-9).all()


@pytest.mark.parametrize("n", [7, 33, 64, 100, 128])
def test_other_sizes_are_bit_identical_to_the_c_oracle(rb, n):
    """Edge sizes of the Sphere family: tiny, odd (padded row stride), the K=2 / K=4 boundary (64), K=4, the maximum
    (128).  Three pairs each under a 12-outer-iteration protocol: x, y and summaries equal the C oracle bit for bit."""
    from oracle.c import binding as detc
    B = 3
    Z, x0, y0 = rb.datagen.nonnegpca_batch(500 + n, B, n)
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=12, inner_maxiter=1000, tolresid=0, maxtime=1e9)
    bs = rb.BatchSolver.nonnegpca_from_arrays(Z, x0, y0)
    bs.set_options(opt, 0, 0)
    x, y, sm, _ = bs.solve()
    bs.close()
    xo, yo, smo = detc.solve_many(Z, x0, y0, {"maxiter": 12, "inner_maxiter": 1000, "tolresid": 0})
    assert np.array_equal(x, xo) and np.array_equal(y, yo)
    assert np.array_equal(sm[:, :15], smo[:, :15])
    assert np.allclose(np.linalg.norm(x, axis=1), 1.0, atol=1e-14) and (x > 0).all()


def test_shared_Z_many_initial_points(rb, datasets):
    """'instances x initial points': one Z (batch_z = 1) shared by 32 feasible starting points (generator.py:46-51
    law).  Same results as 32 separate pairs carrying their own copy of Z; all reach the same KKT quality."""
    d = datasets["NonnegPCA/1"]
    pts = rb.datagen.more_initial_points(d["initx_a"], seed=1, count=32)
    y0 = np.ones((32, 50))
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=30, inner_maxiter=1000, tolresid=0, maxtime=1e9)
    shared = rb.BatchSolver.nonnegpca_from_arrays(d["Z"][None], pts, y0)
    shared.set_options(opt, 0, 0)
    xs, ys, sms, _ = shared.solve()
    shared.close()
    own = rb.BatchSolver.nonnegpca_from_arrays(np.repeat(d["Z"][None], 32, axis=0), pts, y0)
    own.set_options(opt, 0, 0)
    xo, yo, smo, _ = own.solve()
    own.close()
    assert np.array_equal(xs, xo) and np.array_equal(ys, yo) and np.array_equal(sms[:, :15], smo[:, :15])
    assert (sms[:, 1] < 1e-9).all()
    # a nonconvex problem: different starts may end in different KKT points, each a feasible unit vector
    assert np.allclose(np.linalg.norm(xs, axis=1), 1.0, atol=1e-14) and (xs > -1e-12).all()


def test_error_paths_through_the_c_abi(rb):
    """Call-order and argument errors come back as codes + messages (RiptrmError), never as a crash."""
    import ctypes as C
    lib = rb.load_library()
    h = C.c_void_p()
    rb._lib.check(lib.riptrm_create(1, 50, 1, 50, 4, 0, C.byref(h)))
    x = np.ones((4, 50)) / np.sqrt(50)
    with pytest.raises(rb.RiptrmError, match="riptrm_set_<family>"):
        rb._lib.check(lib.riptrm_solve(h, rb._lib.ptr(x), rb._lib.ptr(x), None, None, None, None, 0, None))
    Z = np.zeros((4, 50, 50))
    rb._lib.check(lib.riptrm_set_nonnegpca(h, rb._lib.ptr(Z), 4, C.c_double(0.0), 0))
    with pytest.raises(rb.RiptrmError, match="riptrm_set_options"):
        rb._lib.check(lib.riptrm_solve(h, rb._lib.ptr(x), rb._lib.ptr(x), None, None, None, None, 0, None))
    with pytest.raises(rb.RiptrmError, match="batch_z"):
        rb._lib.check(lib.riptrm_set_nonnegpca(h, rb._lib.ptr(Z), 3, C.c_double(0.0), 0))
    with pytest.raises(rb.RiptrmError):
        rb._lib.check(lib.riptrm_set_rosenbrock(h, C.c_double(1.0), C.c_double(0.01)))
    assert lib.riptrm_destroy(h) == 0
    with pytest.raises(rb.RiptrmError, match="n <= 128"):
        rb._lib.check(lib.riptrm_create(1, 129, 1, 129, 1, 0, C.byref(h)))
    with pytest.raises(ValueError, match="not supported"):        # the reference's own error (RIPTRM.py:453-454)
        rb.RIPTRM({"TRS_solver": "Newton"}).run_batch([None], structures=[rb.NonnegPCAStructure(Z=Z[0], x0=x[0], y0=x[0])])
    # the exact solver keeps three dim x dim matrices per pair in shared memory: Sphere(n) up to n = 64
    big = rb.NonnegPCAStructure(Z=np.eye(80), x0=np.full(80, 80 ** -0.5), y0=np.ones(80))
    with pytest.raises(rb.RiptrmError, match="n <= 64"):
        rb.RIPTRM({"TRS_solver": "Exact_RepMat", "maxiter": 1}).run_batch([None], structures=[big])


def test_tmem_and_shared_memory_kernels_agree_bit_for_bit(rb, monkeypatch):
    """n = 50 runs with S in Tensor Memory by default; RIPTRM_SPHERE_NO_TMEM=1 selects the shared-memory kernel.  Same
    arithmetic, same order: identical x, y, summaries and trace rows (96 pairs, two-launch schedule forced)."""
    Z, x0, y0 = rb.datagen.nonnegpca_sweep(321, 24, 4)
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=20, inner_maxiter=1000, tolresid=0, maxtime=1e9,
               schedule_split=4)
    res = []
    for no_tmem in (False, True):
        if no_tmem:
            monkeypatch.setenv("RIPTRM_SPHERE_NO_TMEM", "1")
        else:
            monkeypatch.delenv("RIPTRM_SPHERE_NO_TMEM", raising=False)
        bs = rb.BatchSolver.nonnegpca_from_arrays(Z, x0, y0)
        bs.set_options(opt, 1, 400)
        x, y, sm, tr = bs.solve()
        bs.close()
        res.append((x, y, sm, tr))
    (xa, ya, sa, ta), (xb, yb, sb, tb) = res
    assert np.array_equal(xa, xb) and np.array_equal(ya, yb) and np.array_equal(sa, sb)
    T = rb._lib.TR
    cols = [i for name, i in T.items() if name != "time"]
    rows = sa[:, rb._lib.SM["trace_rows"]].astype(int)
    for i in range(len(rows)):
        a, b = ta[i, :rows[i]][:, cols], tb[i, :rows[i]][:, cols]
        assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()


def test_dataset_files_in_output_files_out(rb, datasets, tmp_path):
    """The reference's file formats either side of the path: dataset CSVs -> `run` on the loaded structure -> the
    `save_output` file set; same result as the in-memory route, final cost equal to the golden run's to 1e-8."""
    import pandas as pd
    g = load_golden("nonnegpca_1_a_K40")
    d = datasets["NonnegPCA/1"]
    rb.io.save_dataset(str(tmp_path / "dataset" / "NonnegPCA" / "1"), dim=[[50]], Z=d["Z"], initx_a=d["initx_a"],
                       initineqLagmult=d["initineqLagmult"])
    st = rb.io.load_structure("NonnegPCA", str(tmp_path / "dataset"), 1, "a")
    option = {"TRS_solver": "tCG", "second_order_stationarity": False, "tolresid": 0, "maxtime": 1e9, "maxiter": 40}
    out = rb.RIPTRM(option).run(st)
    mem, _ = _solve(rb, rb.NonnegPCAStructure(Z=d["Z"], x0=d["initx_a"], y0=d["initineqLagmult"]), maxiter=40)
    assert np.array_equal(out.x, mem.x) and out.log["cost"] == mem.log["cost"]
    path = rb.io.save_output(out, str(tmp_path / "intermediate" / "NonnegPCA" / "1" / "a"))
    log = pd.read_csv(path + "/RIPTRM_tCG_log.csv")
    assert int((log["inner_status"] == "converged").sum()) == 40
    assert abs(log["cost"].iloc[-1] - g["log"]["cost"][-1]) < REL_TOL * abs(g["log"]["cost"][-1])
    assert np.array_equal(np.loadtxt(path + "/RIPTRM_tCG_x.csv"), out.x)


def test_two_warps_per_copy_of_S_is_bit_identical(rb, monkeypatch):
    """Large batches whose instances come with an even number of initial points run two warps per TMEM copy of S
    (sphere_tmem2_kernel, 16 warps per SM); RIPTRM_SPHERE_NO_SIBLINGS=1 selects the one-warp-per-copy kernel.  Same
    arithmetic per pair: every output must be the same bits, with and without the re-sorting launches."""
    inst, ipp = 1536, 2
    Z, x0, y0 = rb.datagen.nonnegpca_sweep(500, inst, ipp)
    res = {}
    for split in (0, -1, 5):    # automatic (three re-sorts + fast lane), single launch, one re-sort at outer iteration 5
        for sib in ("1", None):
            if sib is None:
                monkeypatch.delenv("RIPTRM_SPHERE_NO_SIBLINGS", raising=False)
            else:
                monkeypatch.setenv("RIPTRM_SPHERE_NO_SIBLINGS", sib)
            opt = rb.options.default_option()
            opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=30, inner_maxiter=1000, tolresid=0, maxtime=1e9,
                       schedule_split=split)
            bs = rb.BatchSolver.nonnegpca_from_arrays(Z, x0, y0)
            bs.set_options(opt, 0, 0)
            res[(split, sib)] = bs.solve()
            bs.close()
    ref = res[(-1, "1")]
    for key, (x, y, sm, _) in res.items():
        assert np.array_equal(x, ref[0]) and np.array_equal(y, ref[1]) and np.array_equal(sm[:, :15], ref[2][:, :15]), key
    assert (ref[2][:, rb._lib.SM["residual"]] < 1e-8).all()


@pytest.mark.gpu
def test_log_that_overflows_its_buffer_resolves_only_those_pairs(rb, datasets, monkeypatch):
    """run_batch sizes the per-inner-iteration log buffer from maxiter; a pair that needs more rows is solved again (alone, the
    solve is deterministic) with room for every row, the rest of the batch is not: logs identical to the roomy run."""
    import riptrm_b200.solver as solver_mod
    d = datasets["NonnegPCA/1"]
    rs = np.random.RandomState(3)
    starts = [d["initx_a"]] + [v / np.linalg.norm(v) for v in (rs.rand(d["initx_a"].size) + 0.05 for _ in range(2))]
    sts = [rb.NonnegPCAStructure(Z=d["Z"], x0=x0, y0=d["initineqLagmult"]) for x0 in starts]
    opt = {"TRS_solver": "tCG", "second_order_stationarity": False, "tolresid": 0, "maxtime": 1e9, "maxiter": 12}
    roomy = rb.RIPTRM(opt).run_batch([None] * 3, structures=sts)
    rows = sorted(len(o.log["iteration"]) for o in roomy)
    monkeypatch.setattr(solver_mod, "TRACE_CAPACITY_OVERRIDE", (rows[0] + rows[1]) // 2 if rows[0] < rows[1] else rows[0] - 1)
    tight = rb.RIPTRM(opt).run_batch([None] * 3, structures=sts)
    for a, b in zip(roomy, tight):
        assert len(a.log["iteration"]) == len(b.log["iteration"])
        for k in a.log:
            if k == "time":
                continue
            assert all((u == v) or (u != u and v != v) for u, v in zip(a.log[k], b.log[k])), k
        assert np.array_equal(a.x, b.x) and np.array_equal(a.ineqLagmult, b.ineqLagmult)
