"""SURVEY section 8f rank 4 on the GPU: the condensed Newton system of the reference's Riemannian interior-point method,
`OperatorAw = OperatorHessLag + OperatorTHETA` (src/solver/RIPM.py:484-511), solved by `RepresentMatMethod` (:238-300) and by
`TangentSpaceConjResMethod` (src/solver/utils.py:582-618) -- `riptrm_newton` against the NumPy restatement
(oracle.riptrm_oracle.newton_repmat / conj_res on the per-constraint operators)."""
import numpy as np
import pytest

from helpers import nonnegpca_problem, rosenbrock_problem, stableid_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import riptrm_b200
    return riptrm_b200


def _case(rb, datasets, which):
    if which == "NonnegPCA":
        P = nonnegpca_problem(datasets)
        return rb.NonnegPCAStructure(Z=P.Z, x0=P.initialpoint, y0=P.initialineqLagmult), P
    if which == "Rosenbrock":
        P = rosenbrock_problem()
        return rb.RosenbrockStructure(n=5, k=3, alpha=1e7, x0=P.initialpoint, y0=P.initialineqLagmult), P
    P = stableid_problem(datasets, "a")
    conspec = np.array([[k, r, c, a, b] for (k, r, c, a, b) in P.spec], dtype=float)
    return rb.StableIdStructure(X=P.X, XP=P.XP, h=P.h, conspec=conspec, x0=P.initialpoint, y0=P.initialineqLagmult), P


def _flat(st, v):
    return st.pack_x(v)[None]


@pytest.mark.parametrize("which", ["NonnegPCA", "Rosenbrock", "StableIdentification"])
def test_newton_system_matches_oracle(rb, datasets, which):
    from oracle import riptrm_oracle as O
    from riptrm_b200.basis import deterministic_basisfun
    rng = np.random.RandomState(21)
    st, P = _case(rb, datasets, which)
    man, x = P.manifold, P.initialpoint
    m = len(P.initialineqLagmult)
    z = 0.5 + rng.rand(m)                      # multipliers and slacks of an interior-point iterate: independent of s(x)
    s = 0.2 + rng.rand(m)
    amb = [rng.randn(*np.shape(a)) for a in x] if isinstance(x, list) else rng.randn(*np.shape(x))
    c = man.projection(x, amb) if hasattr(man, "projection") else man.to_tangent_space(x, amb)
    Aw = O.newton_operator(P, x, z, s)
    bs = rb.BatchSolver([st])
    # ---- RepresentMatMethod
    ref, Amat = O.newton_repmat(P, x, z, s, c, deterministic_basisfun(man, x))
    dx, info = bs.newton(bs.x0, z[None], s[None], _flat(st, c), "RepMat")
    got = st.unpack_x(dx[0])
    scale = man.norm(x, ref)
    parts = zip(got, ref) if isinstance(got, list) else [(got, ref)]
    cond = np.linalg.cond(Amat)
    for g, r in parts:
        assert np.max(np.abs(np.asarray(g) - np.asarray(r))) <= 1e-13 * cond * scale + 1e-12 * scale, (which, cond)
    assert info[0, 0] == 0 and info[0, 1] <= 1e-13 * cond
    assert abs(info[0, 2] - scale) <= (1e-13 * cond + 1e-10) * scale
    eig = np.linalg.eigvalsh(Amat)
    assert abs(info[0, 3] - eig[0]) <= 1e-10 * max(abs(eig[0]), abs(eig[-1]))
    # ---- TangentSpaceConjResMethod, the reference's defaults (KrylovTolrelresid 1e-9, KrylovMaxIteration 1000) and a cap
    norm_c = man.norm(x, c)
    for tol, cap in ((1e-9, 1000), (1e-6, 1000), (1e-3, 1000), (0.0, 7)):
        v, t, rel = O.conj_res(man, x, Aw, c, tol, cap)
        dx, info = bs.newton(bs.x0, z[None], s[None], _flat(st, c), "Krylov", tol=tol, maxiter=cap)
        got = st.unpack_x(dx[0])
        it = int(info[0, 0])
        # Iteration counts: identical on the well-conditioned sphere system (cond 2.3, 13 iterations) and under the cap; on
        # the indefinite systems of the other two workloads (cond 1e7 / 1.6e4, 28 / 130 iterations) conjugacy is lost to
        # rounding and two correct runs stop within ~15 % of each other (observed 141 vs 130)
        if which == "NonnegPCA" or tol == 0.0:
            assert it == t, (which, tol, it, t)
            parts = zip(got, v) if isinstance(got, list) else [(got, v)]
            for g, r in parts:
                assert np.max(np.abs(np.asarray(g) - np.asarray(r))) <= 1e-6 * man.norm(x, v), (which, tol)
            assert abs(info[0, 1] - rel) <= 1e-6 * max(rel, 1e-12) + 1e-14
        elif tol >= 1e-6:
            assert abs(it - t) <= max(3, 0.2 * t), (which, tol, it, t)
        else:
            # at the reference's default 1e-9 these systems sit close to the attainable accuracy (eps * cond = 1e-9 / 1e-12):
            # the residual curve plateaus and the iteration at which it dips under the tolerance depends on the last bits
            # (observed 130 / 141 / 337 on the Product system for three arithmetic variants) -- only the result is checked
            assert it <= cap
        if tol > 0:
            # whatever the count, the stopping rule holds for the returned point (checked with the oracle's operator) and it
            # agrees with the direct solve up to tol * cond
            gv = O._amb(man, got) if isinstance(got, list) else got
            res = Aw(gv)
            diff = [a - b for a, b in zip(c, res)] if isinstance(got, list) else c - res
            true_rel = man.norm(x, O._amb(man, diff) if isinstance(got, list) else diff) / norm_c
            assert info[0, 1] < tol and true_rel <= 1.5 * tol + 1e-12 * cond, (which, tol, info[0, 1], true_rel)
            parts = zip(got, ref) if isinstance(got, list) else [(got, ref)]
            for g, r in parts:
                assert np.max(np.abs(np.asarray(g) - np.asarray(r))) <= 10 * tol * cond * scale
    bs.close()
