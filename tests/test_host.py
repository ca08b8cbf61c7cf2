"""Host-side logic that needs no GPU: option defaults / schedules, log reconstruction from trace rows,
the synthetic-instance generator, structure recognition of reference-style closures."""
import math
import os

import numpy as np
import pytest

import riptrm_b200 as rb
from riptrm_b200 import _lib, options


def test_defaults_match_oracle_defaults():
    from oracle.riptrm_oracle import default_option as oracle_defaults
    d, o = options.default_option(), oracle_defaults()
    for k, v in o.items():
        if callable(v):
            for mu in (0.1, 1e-6, 1e-16):
                if k.startswith("forcing"):
                    assert d[k](mu) == v(mu)
        elif k not in ("TRS_solver", "second_order_stationarity", "basisfun"):
            assert d[k] == v, k
    # the reference's class defaults (RIPTRM.py:320-321): Exact_RepMat + second order; configs override to tCG
    assert d["TRS_solver"] == "Exact_RepMat" and d["second_order_stationarity"] is True


def test_barrier_schedule():
    opt = options.default_option()
    opt["maxiter"] = 30
    mu, tolL, tolC = options.barrier_schedule(opt)
    assert len(mu) == 31 and mu[0] == 0.1
    m = 0.1
    for k in range(30):
        m = max(1e-15, 0.5 * m ** 1.01)       # RIPTRM.py:890-891
        assert mu[k + 1] == m
    assert tolL[0] == 0.1 and tolC[0] == 1e-4
    assert tolL[-1] == max(mu[-1], 1e-14)
    opt["do_simple_barrier_parameter_update"] = False
    mu2, _, _ = options.barrier_schedule(opt)
    assert mu2[1] == min(0.8 * 0.1, 0.5 * 0.1 ** 1.01)


def test_trs_solver_options():
    """Both trust-region solvers of the reference are accepted (the class default is 'Exact_RepMat' with the second-order
    test, RIPTRM.py:323-324); an unknown solver raises the reference's error (:453-454)."""
    opt = options.default_option()
    assert opt["TRS_solver"] == "Exact_RepMat" and opt["second_order_stationarity"] is True and callable(opt["basisfun"])
    options.check_supported(opt)
    o, keep = options.to_c_options(dict(opt, maxiter=5), 1, 64)
    assert (o.trs_solver, o.second_order_stationarity, o.trs_tolhardcase) == (1, 1, 1e-8)
    assert np.allclose(keep[3], keep[0])                 # forcing_function_second_order(mu) = mu
    o, keep = options.to_c_options(dict(opt, maxiter=5, TRS_solver="tCG"), 1, 64)
    assert (o.trs_solver, o.second_order_stationarity) == (0, 0) and keep[3] is None   # no eigenvalue test under tCG (:599)
    with pytest.raises(ValueError, match="not supported"):
        options.check_supported(dict(opt, TRS_solver="Newton"))
    with pytest.raises(NotImplementedError):
        options.check_supported(dict(opt, checkTRSoptimality=True))


def test_deterministic_tangent_bases_are_orthonormal(datasets):
    """riptrm_b200/basis.py (`basisfun` for reproducible reference runs; the kernels build the same bases): dim vectors,
    tangent, orthonormal in the manifold's metric."""
    from oracle import manifolds as M
    from riptrm_b200.basis import deterministic_basisfun
    rng = np.random.RandomState(0)
    d = datasets["StableIdentification/1"]
    cases = [(M.Sphere(50), datasets["NonnegPCA/1"]["initx_a"]),
             (M.Grassmann(5, 3), np.linalg.qr(rng.randn(5, 3))[0]),
             (M.Product([M.SkewSymmetric(5), M.SymmetricPositiveDefinite(5), M.SymmetricPositiveDefinite(5)]),
              [d["initJ_a"], d["initR_a"], d["initQ_a"]])]
    for man, x in cases:
        B = deterministic_basisfun(man, x)
        assert len(B) == man.dim
        G = np.array([[man.inner_product(x, a, b) for b in B] for a in B])
        assert np.max(np.abs(G - np.eye(man.dim))) < 1e-12
        for b in B[:: max(1, man.dim // 7)]:
            pb = man.to_tangent_space(x, b)
            diff = (pb - b) if not isinstance(b, list) else [u - v for u, v in zip(pb, b)]
            assert man.norm(x, diff) < 1e-12


def test_trace_to_log_schema():
    nan = float("nan")
    T = _lib.TR
    row0 = np.full(_lib.TRACE_FIELDS, nan)
    row0[T["iteration"]] = 0
    row0[T["mu"]] = 0.1
    row0[T["cost"]] = -0.5
    row0[T["maxabsLagmult"]] = 1.0
    row1 = np.zeros(_lib.TRACE_FIELDS)
    row1[T["iteration"]] = 1
    row1[T["num_inner"]] = 1
    row1[T["dxtype"]] = 2
    row1[T["inner_status"]] = 3
    row1[T["radius_update"]] = 2
    row1[T["dual_clipping"]] = 1
    row1[T["tcg_iters"]] = 7
    row1[T["mineigvalHw"]] = nan
    row2 = row1.copy()                       # a row of the exact trust-region solver
    row2[T["dxtype"]] = 7
    row2[T["tcg_iters"]] = nan
    row2[T["mineigvalHw"]] = 2.5
    logx = rb.trace_to_log(np.stack([row0, row2]))
    assert logx["dxtype"] == [None, "interior"] and logx["mineigvalHw"] == [None, 2.5] and logx["tcg_iters"] == [None, None]
    assert rb.trace_to_log(np.stack([row0[:25], row1[:25]]))["mineigvalHw"] == [None, None]   # 25-field rows (C oracle)
    log = rb.trace_to_log(np.stack([row0, row1]))
    # reference column order: base_solver.py:58-76, utils.py:356-364, RIPTRM.py:980-1024
    assert list(log)[:11] == ["iteration", "time", "cost", "distance", "residual", "gradnorm", "complviolation",
                              "dualviolation", "manviolation", "maxviolation", "meanviolation"]
    assert log["inner_status"] == [None, "successful"]
    assert log["dxtype"] == [None, "tCG_EXCEEDED_TR"]
    assert log["radius_update"] == [None, "expanded"]
    assert log["dual_clipping"] == [None, True]
    assert log["num_inner"] == [None, 1] and log["tcg_iters"] == [None, 7]
    assert log["mineigvalHw"] == [None, None]
    assert len({len(v) for v in log.values()}) == 1


def test_datagen_follows_the_oracle_generator():
    from oracle.problems import nonnegpca_generate_instance
    for seed in (0, 5, 4095):
        Z, x0, y0 = rb.datagen.nonnegpca_instance(50, seed=seed)
        Zo, xo, yo = nonnegpca_generate_instance(50, seed=seed)
        assert np.array_equal(Z, Zo) and np.array_equal(x0, xo) and np.array_equal(y0, yo)
        assert np.all(x0 > 0) and abs(np.linalg.norm(x0) - 1) < 1e-15
        assert np.max(np.abs(Z - Z.T)) > 0.1      # Z is not symmetric
    Zb, xb, yb = rb.datagen.nonnegpca_batch(3, 2, 50)
    assert np.array_equal(Zb[1], rb.datagen.nonnegpca_instance(50, seed=4)[0])


def test_structure_recognition_from_closures(datasets):
    """A reference-style problem (pymanopt-like manifold object + closures over Z / idx) is recognised
    from its closures; unknown problems raise (no CPU fallback)."""
    d = datasets["NonnegPCA/1"]
    Z = d["Z"]

    class Sphere:  # stands for pymanopt.manifolds.Sphere (recognised by class name)
        pass

    def cost(point):
        return -point @ Z @ point

    def make(idx):
        def ineq(point):
            return -point[idx]
        return ineq

    class P:
        manifold = Sphere()
        initialpoint = d["initx_a"]
        initialineqLagmult = d["initineqLagmult"]
        ineqconstraints_all = [make(i) for i in range(50)]
    P.cost = staticmethod(cost)
    st = rb.structure_from_problem(P)
    assert isinstance(st, rb.NonnegPCAStructure) and st.shape == (50, 1, 50) and st.Z is not None
    assert np.array_equal(st.Z, Z)

    class Q(P):
        manifold = type("Euclidean", (), {})()
    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        rb.structure_from_problem(Q)


def test_stableid_conspec_expansion(datasets):
    cs = rb.StableIdStructure.conspec_from_constset(datasets["StableIdentification/1"]["constset"])
    assert cs.shape == (16, 5)
    assert sorted(set(cs[:, 0])) == [0.0, 1.0, 2.0]


def test_sweep_generator_matches_oracle_and_scales_to_8_gpus():
    """Instance ids of rank 7 of an 8-GPU sweep (28672..32767) must still give valid NumPy seeds; product and oracle
    generators agree; the initial points of an instance are distinct feasible unit vectors sharing its Z."""
    from oracle.problems import nonnegpca_generate_sweep
    for first in (0, 7 * 4096 + 4000, 10 ** 6):
        a = rb.datagen.nonnegpca_sweep(first, 3, 4)
        b = nonnegpca_generate_sweep(first, 3, 4)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        Z, x0, y0 = a
        assert Z.shape == (3, 50, 50) and x0.shape == (12, 50)
        assert np.allclose(np.linalg.norm(x0, axis=1), 1.0) and (x0 > 0).all()
        assert len({tuple(r) for r in x0}) == 12
        assert np.array_equal(x0[4], rb.datagen.nonnegpca_instance(50, seed=first + 1)[1])
    with pytest.raises(ValueError):
        rb.datagen.nonnegpca_sweep(0, 2, 17)


def test_dataset_csv_roundtrip_and_loader(tmp_path, datasets):
    """The reference's dataset format (np.savetxt '%.18e' files under dataset/<problem>/<instance>/) round-trips
    exactly and loads into the structured problem descriptions."""
    d = datasets["NonnegPCA/1"]
    root = tmp_path / "dataset"
    rb.io.save_dataset(str(root / "NonnegPCA" / "7"), dim=[[50]], Z=d["Z"], initx_a=d["initx_a"],
                       initineqLagmult=d["initineqLagmult"])
    st = rb.io.load_structure("NonnegPCA", str(root), instance=7, initialpoint="a")
    assert np.array_equal(st.Z, d["Z"]) and np.array_equal(st.x0, d["initx_a"]) and st.shape == (50, 1, 50)
    s = datasets["StableIdentification/1"]
    sid = str(root / "StableIdentification" / "1")
    rb.io.save_dataset(sid, constset=s["constset"], initineqLagmult=s["initineqLagmult"],
                       **{f"noisyX_{k}": s[f"noisyX_{k}"] for k in range(1, 6)},
                       **{f"init{c}_b": s[f"init{c}_b"] for c in "JRQ"})
    st = rb.io.load_structure("StableIdentification", str(root), instance=1, initialpoint="b")
    assert st.X.shape == (5, 95) and st.XP.shape == (5, 95) and st.conspec.shape == (16, 5)
    assert np.array_equal(st.X[:, :19], s["noisyX_1"][:, :-1]) and np.array_equal(st.XP[:, :19], s["noisyX_1"][:, 1:])
    assert np.array_equal(st.x0[1], s["initR_b"])
    ros = rb.io.load_structure("Rosenbrock", str(root))
    assert ros.shape == (5, 3, 15) and ros.alpha == 1e7
    with pytest.raises(NotImplementedError):
        rb.io.load_structure("Unknown", str(root))


def test_save_output_writes_the_reference_file_set(tmp_path):
    import pandas as pd
    nan = float("nan")
    T = _lib.TR
    rows = np.full((3, _lib.TRACE_FIELDS), nan)
    rows[:, T["iteration"]] = [0, 1, 1]
    rows[:, T["cost"]] = [-0.5, -0.7, -0.9]
    rows[1:, T["inner_status"]] = [3, 1]
    rows[:, T["maxabsLagmult"]] = 1.0
    opt = options.default_option()
    opt["stoppingcriterion"] = "Max iteration count reached; maxiter=1 after 0.00 seconds"
    out = rb.Output(name="RIPTRM_tCG", x=np.arange(3.0), option=opt, log=rb.trace_to_log(rows),
                    ineqLagmult=np.ones(3), eqLagmult=[])
    path = rb.io.save_output(out, str(tmp_path / "intermediate" / "NonnegPCA" / "1" / "a"))
    files = sorted(os.listdir(path))
    assert files == [f"RIPTRM_tCG_{a}.csv" for a in sorted(("name", "x", "option", "log", "ineqLagmult", "eqLagmult"))]
    log = pd.read_csv(os.path.join(path, "RIPTRM_tCG_log.csv"))
    assert list(log["cost"]) == [-0.5, -0.7, -0.9] and list(log["inner_status"])[1:] == ["successful", "converged"]
    assert np.array_equal(np.loadtxt(os.path.join(path, "RIPTRM_tCG_x.csv")), np.arange(3.0))
    optcsv = pd.read_csv(os.path.join(path, "RIPTRM_tCG_option.csv"))
    assert len(optcsv) == 1 and optcsv["maxiter"][0] == 100


def test_save_output_product_point_is_block_file(tmp_path):
    """For a Product-manifold point (list of matrices) the reference's writer falls into csv.writer.writerows: one line per
    component, one cell per matrix row holding numpy's str() of that row (8 significant digits).  The strict-complementarity
    analyzer parses exactly that (src/StableIdentification/analyzer_strict_complementarity.py:6-34): strip the brackets,
    split on blanks, stack."""
    import csv
    rs = np.random.RandomState(3)
    x = [rs.randn(5, 5) for _ in range(3)]
    out = rb.Output(name="RIPTRM_tCG", x=x, option={"maxiter": 1}, log={"iteration": [0]}, ineqLagmult=np.ones(16), eqLagmult=[])
    path = rb.io.save_output(out, str(tmp_path))
    mats = []
    with open(os.path.join(path, "RIPTRM_tCG_x.csv"), newline="") as f:
        for row in csv.reader(f):
            vecs = [np.array(cell.strip().lstrip("[").rstrip("]").split(), dtype=float) for cell in row if cell.strip()]
            mats.append(np.vstack(vecs))
    got = np.stack(mats)
    assert got.shape == (3, 5, 5) and np.max(np.abs(got - np.stack(x))) < 1e-7
    assert np.loadtxt(os.path.join(path, "RIPTRM_tCG_ineqLagmult.csv")).shape == (16,)
    assert os.path.getsize(os.path.join(path, "RIPTRM_tCG_eqLagmult.csv")) == 0


def test_stableid_sweep_points_are_strictly_feasible(datasets):
    """More starting points for the StableIdentification sweep: the reference's own 20 come first unchanged, the rest
    are perturbations that stay skew / symmetric positive definite and strictly inside every constraint."""
    d = datasets["StableIdentification/1"]
    conspec = rb.StableIdStructure.conspec_from_constset(d["constset"])
    base = [[d[f"init{c}_{pt}"] for c in "JRQ"] for pt in "abcdefghijklmnopqrst"]
    pts = rb.datagen.stableid_more_initial_points(base, conspec, 200, seed=5)
    assert len(pts) == 200 and all(np.array_equal(pts[i][k], base[i][k]) for i in range(20) for k in range(3))
    for J, R, Q in pts:
        assert np.array_equal(J, -J.T) and np.array_equal(R, R.T) and np.array_equal(Q, Q.T)
        assert np.linalg.eigvalsh(R).min() > 0 and np.linalg.eigvalsh(Q).min() > 0
        assert rb.datagen.stableid_constraint_values((J - R) @ Q, conspec).max() < 0
    assert not np.array_equal(pts[20][1], base[0][1])
    again = rb.datagen.stableid_more_initial_points(base, conspec, 200, seed=5)
    assert all(np.array_equal(a[k], b[k]) for a, b in zip(pts, again) for k in range(3))


def test_batch_of_different_small_instances_is_rejected(datasets):
    """a Rosenbrock / StableIdentification handle binds ONE instance's data, so a batch mixing instances must
    raise instead of being solved against the first instance's data."""
    from riptrm_b200.solver import validate_batch
    x0 = np.eye(5)[:, :3]
    a = rb.RosenbrockStructure(n=5, k=3, alpha=1e7, x0=x0, y0=np.ones(15))
    b = rb.RosenbrockStructure(n=5, k=3, alpha=1e6, x0=x0, y0=np.ones(15))
    validate_batch([a, a])
    with pytest.raises(ValueError, match="one instance"):
        validate_batch([a, b])
    d = datasets["StableIdentification/1"]
    Xs = [d[f"noisyX_{k}"] for k in range(1, 6)]
    X, XP = np.hstack([x[:, :-1] for x in Xs]), np.hstack([x[:, 1:] for x in Xs])
    cs = rb.StableIdStructure.conspec_from_constset(d["constset"])
    pt = [d[f"init{c}_a"] for c in "JRQ"]
    s1 = rb.StableIdStructure(X=X, XP=XP, h=0.02, conspec=cs, x0=pt, y0=d["initineqLagmult"])
    s2 = rb.StableIdStructure(X=X.copy(), XP=XP.copy(), h=0.02, conspec=cs.copy(), x0=pt, y0=d["initineqLagmult"])
    s3 = rb.StableIdStructure(X=X * 1.5, XP=XP, h=0.02, conspec=cs, x0=pt, y0=d["initineqLagmult"])
    validate_batch([s1, s2])          # equal data in different arrays is one instance
    with pytest.raises(ValueError, match="one instance"):
        validate_batch([s1, s3])
    with pytest.raises(ValueError, match="one family"):
        validate_batch([a, s1])


def test_user_manviofun_and_callbackfun_are_checked_loudly(datasets):
    """VERDICT r1 missing #6: a caller's manviofun / callbackfun must not be dropped silently."""
    import types
    import warnings
    from riptrm_b200.solver import check_user_functions, builtin_manvio
    d = datasets["NonnegPCA/1"]
    st = rb.NonnegPCAStructure(Z=d["Z"], x0=d["initx_a"], y0=d["initineqLagmult"])
    problem = types.SimpleNamespace(initialpoint=d["initx_a"], initialineqLagmult=d["initineqLagmult"])
    log = {c: [0.5] for c in ("iteration", "cost", "distance", "residual", "gradnorm", "complviolation", "dualviolation",
                              "manviolation", "maxviolation", "meanviolation")}
    opt = options.default_option()
    assert check_user_functions(opt, problem, st, log) == []                        # class defaults
    opt["manviofun"] = lambda problem, x: np.linalg.norm(x) - 1                      # the reference simulator's
    assert check_user_functions(opt, problem, st, log) == []
    opt["manviofun"] = lambda problem, x: abs(x[0])                                  # something else: refuse
    with pytest.raises(NotImplementedError, match="manviofun"):
        check_user_functions(opt, problem, st, log)
    opt["manviofun"] = lambda problem, x: 0
    opt["callbackfun"] = lambda problem, x, y, z, ev: dict(ev, second_order_residual=1.0)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert check_user_functions(opt, problem, st, log) == ["second_order_residual"]
    assert any("callbackfun" in str(x.message) for x in w)
    assert builtin_manvio(st, 2 * d["initx_a"]) == pytest.approx(1.0)


def test_package_import_does_not_need_torch():
    """the library / drop-in module import without torch (bench.py and `sharding` are its only users)."""
    import subprocess
    import sys
    code = ("import sys; sys.modules['torch'] = None\n"
            f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
            "import riptrm_b200 as rb; rb.RIPTRM({'TRS_solver': 'tCG'}); print('ok')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr[-1000:]


def test_standin_coordinator_problems_are_recognised(datasets):
    """tests/dropin_standins.py (the caller side of the drop-in as the reference's coordinators write it, used by the GPU
    drop-in test) yields problems whose structure is recovered from the closures alone."""
    import dropin_standins as D
    st = rb.structure_from_problem(D.build_problem("NonnegPCA", datasets))
    assert isinstance(st, rb.NonnegPCAStructure) and np.array_equal(st.Z, datasets["NonnegPCA/1"]["Z"])
    st = rb.structure_from_problem(D.build_problem("Rosenbrock", datasets))
    assert isinstance(st, rb.RosenbrockStructure) and (st.alpha, st.offset, st.shape) == (1e7, 0.01, (5, 3, 15))
    st = rb.structure_from_problem(D.build_problem("StableIdentification", datasets, "t"))
    cs = rb.StableIdStructure.conspec_from_constset(datasets["StableIdentification/1"]["constset"])
    assert isinstance(st, rb.StableIdStructure) and np.array_equal(st.conspec, cs) and st.X.shape == (5, 95)
    assert np.array_equal(st.x0[2], datasets["StableIdentification/1"]["initQ_t"])


def test_bench_arms_describe_the_same_workload():
    """The driver compares the `config` of `bench.py` and `bench.py --impl reference`: both come from one function of the
    arguments, steps in flight included."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_bench_mod", os.path.join(root, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
        args = mod.parse()
    finally:
        sys.argv = argv
    assert args.in_flight == 2 and args.gpus == 1 and args.scaling == "weak"
    c1, c8 = mod.config_dict(args, 1), mod.config_dict(args, 8)
    assert c1["in_flight"] == 2 and c1["pairs_per_gpu"] == 16384 and c8["pairs_total"] == 8 * 16384
    assert "flushed" in c1["l2"] and mod.UNIT == "pairs/s"
    args.scaling = "strong"
    assert mod.config_dict(args, 8)["pairs_per_gpu"] == 2048 and mod.instances_per_rank(args, 8) == 512


def test_recognition_with_the_attribute_layout_of_real_pymanopt(datasets, monkeypatch):
    """pymanopt 2.2's `Function` keeps the user's function as `_original_function`, a backend object as `_backend` and the
    backend's `prepare_function(function)` result as `_function` (for the autograd backend the function itself, for others a
    wrapper).  The closure walk does not depend on attribute names: a Function laid out that way -- with `_function` a
    functools.wraps wrapper, the harder case -- is recognised like the stand-in's."""
    import functools

    import dropin_standins as D
    import pymanopt.function as pf

    class _Backend:
        def __init__(self):
            self.name = "Autograd"

        def prepare_function(self, function):
            @functools.wraps(function)
            def prepared(*args):
                return function(*args)
            return prepared

    class RealLayoutFunction(pf.Function):
        def __init__(self, function, manifold):
            super().__init__(function, manifold)
            self._original_function = function
            self._backend = _Backend()
            self._function = self._backend.prepare_function(function)
            self._gradient = None
            self._hessian = None

    monkeypatch.setattr(pf, "Function", RealLayoutFunction)
    pb = D.build_problem("NonnegPCA", datasets)
    assert type(pb.cost).__name__ == "RealLayoutFunction" and hasattr(pb.cost, "_original_function")
    st = rb.structure_from_problem(pb)
    assert isinstance(st, rb.NonnegPCAStructure) and np.array_equal(st.Z, datasets["NonnegPCA/1"]["Z"])
    st = rb.structure_from_problem(D.build_problem("Rosenbrock", datasets))
    assert isinstance(st, rb.RosenbrockStructure) and (st.alpha, st.offset, st.shape) == (1e7, 0.01, (5, 3, 15))
    st = rb.structure_from_problem(D.build_problem("StableIdentification", datasets, "t"))
    assert isinstance(st, rb.StableIdStructure) and st.X.shape == (5, 95)
