"""Full BASELINE sizes on the GPU, checked through size-independent properties plus sampled bit-exact comparison:
config 5 (4096 instances x 4 initial points) and config 4 (n = 20000, p = 10)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import riptrm_b200
    return riptrm_b200


def test_config5_full_sweep_invariants_and_sampled_bit_identity(rb):
    from oracle.c import binding as detc
    I, ipp = 4096, 4
    Z, x0, y0 = rb.datagen.nonnegpca_sweep(0, I, ipp)
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=30, inner_maxiter=1000, tolresid=0, maxtime=1e9)
    bs = rb.BatchSolver.nonnegpca_from_arrays(Z, x0, y0)
    bs.set_options(opt, 0, 0)
    x, y, sm, _ = bs.solve()
    bs.close()
    SM = rb._lib.SM
    B = I * ipp
    assert x.shape == (B, 50) and np.isfinite(x).all() and np.isfinite(y).all() and np.isfinite(sm).all()
    assert (sm[:, SM["stop_reason"]] == 2).all() and (sm[:, SM["outer_iters"]] == 30).all()
    assert sm[:, SM["residual"]].max() < 1e-9                      # every pair reaches the KKT level of the protocol
    assert np.abs(np.linalg.norm(x, axis=1) - 1).max() < 1e-14      # on the sphere
    assert x.min() > 0 and y.min() > 0                              # strictly feasible primal and dual iterates
    mu_end = sm[0, SM["mu"]]
    assert np.abs(x * y - mu_end).max() < 1e-9                      # complementarity y_i s_i -> mu
    cost = -np.einsum("bi,bij,bj->b", x, np.repeat(Z, ipp, axis=0), x)
    assert np.abs(cost - sm[:, SM["cost"]]).max() < 1e-12           # reported objective is the objective of x
    # initial points of one instance may reach different KKT points (nonconvex), but never a worse one than x0
    cost0 = -np.einsum("bi,bij,bj->b", x0, np.repeat(Z, ipp, axis=0), x0)
    assert (cost < cost0).all()
    # sampled pairs against the deterministic C oracle, bit for bit
    rs = np.random.RandomState(0)
    idx = np.sort(rs.choice(B, 96, replace=False))
    xo, yo, smo = detc.solve_many(np.repeat(Z, ipp, axis=0)[idx], x0[idx], y0[idx],
                                  {"maxiter": 30, "inner_maxiter": 1000, "tolresid": 0}, threads=8)
    assert np.array_equal(x[idx], xo) and np.array_equal(y[idx], yo) and np.array_equal(sm[idx, :15], smo[:, :15])


def test_config4_full_size_hessvec_properties(rb):
    """n = 20000, p = 10: Hw against an independent fp64 evaluation on the device (torch), linearity, self-adjointness
    on the tangent space, determinism, and the lock-step tCG's trust-region / tangency invariants."""
    import torch
    n, p = 20000, 10
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    Z = torch.randn((n, n), generator=g, dtype=torch.float64, device=dev) / np.sqrt(n)
    X = torch.rand((n, p), generator=g, dtype=torch.float64, device=dev)
    X = (X / X.norm(dim=0, keepdim=True)).contiguous()
    Y = (0.5 + torch.rand((n, p), generator=g, dtype=torch.float64, device=dev)).contiguous()
    proj = lambda V: V - X * (X * V).sum(dim=0, keepdim=True)
    U = proj(torch.randn((n, p), generator=g, dtype=torch.float64, device=dev)).contiguous()
    V = proj(torch.randn((n, p), generator=g, dtype=torch.float64, device=dev)).contiguous()
    cs = rb.ColumnsSolver(Z, p)
    mu = 0.05
    HU = cs.hessvec(X, Y, mu, U)
    HV = cs.hessvec(X, Y, mu, V)
    # independent evaluation of SURVEY App. A.1 with torch fp64 matmuls
    SV = Z @ V + Z.T @ V
    SX = Z @ X + Z.T @ X
    kappa = (X * SX).sum(dim=0, keepdim=True) + (Y * X).sum(dim=0, keepdim=True)
    ref = proj(-SV) + kappa * V + proj((Y / X) * proj(V))
    assert float((HV - ref).abs().max()) < 1e-9 * float(ref.abs().max())
    HUV = cs.hessvec(X, Y, mu, (U + 2.0 * V).contiguous())
    assert float((HUV - (HU + 2.0 * HV)).abs().max()) < 1e-11 * float(HUV.abs().max())
    uhv, vhu = (U * HV).sum(dim=0), (V * HU).sum(dim=0)
    assert float((uhv - vhu).abs().max()) < 1e-9 * float(uhv.abs().max())
    assert torch.equal(cs.hessvec(X, Y, mu, U), HU)
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=1)
    cs.set_options(opt)
    Delta = 0.01
    eta, info = cs.tcg(X, Y, mu, Delta)
    assert float(eta.norm(dim=0).max()) <= Delta * (1 + 1e-12)
    assert float((X * eta).sum(dim=0).abs().max()) < 1e-12
    assert torch.allclose(eta.norm(dim=0), info[:, 2], rtol=1e-12, atol=0)
    assert int(info[:, 0].min()) >= 1
    cs.close()


def test_config4_full_size_on_stiefel(rb):
    """n = 20000, p = 10 as ONE problem on Stiefel(n, p) (family STIEFEL): Hw against an independent fp64 evaluation on the
    device (torch matmuls), tangency of the product, the tCG's trust-region / tangency invariants, and a
    capped whole solve (maxiter = 2, inner_maxiter = 5) that must stay on the manifold and strictly feasible."""
    import torch
    n, p, eps = 20000, 10, 0.01
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    Z = torch.randn((n, n), generator=g, dtype=torch.float64, device=dev) / np.sqrt(n)
    X = torch.zeros((n, p), dtype=torch.float64, device=dev)
    b = n // p
    for c in range(p):
        X[c * b:(c + 1) * b, c] = torch.rand(b, generator=g, dtype=torch.float64, device=dev) + 0.1
    X = (X / X.norm(dim=0, keepdim=True)).contiguous()
    Y = (0.5 + torch.rand((n, p), generator=g, dtype=torch.float64, device=dev)).contiguous()
    sym = lambda M: 0.5 * (M + M.T)
    proj = lambda V: V - X @ sym(X.T @ V)
    V = proj(torch.randn((n, p), generator=g, dtype=torch.float64, device=dev)).contiguous()
    ss = rb.StiefelSolver(Z, p, eps=eps)
    mu = 0.05
    HV = ss.hessvec(X, Y, mu, V)
    SV, SX = Z @ V + Z.T @ V, Z @ X + Z.T @ X
    C1 = sym(X.T @ SX) + sym(X.T @ Y)
    ref = proj(-SV + V @ C1 + (Y / (X + eps)) * proj(V))
    assert float((HV - ref).abs().max()) < 1e-9 * float(ref.abs().max())
    assert float((X.T @ HV + HV.T @ X).abs().max()) < 1e-9 * float(ref.abs().max())
    assert torch.equal(ss.hessvec(X, Y, mu, V), HV)
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=1)
    ss.set_options(opt)
    Delta = 0.05
    eta, info = ss.tcg(X, Y, mu, Delta)
    assert float(eta.norm()) <= Delta * (1 + 1e-12) and abs(float(eta.norm()) - float(info[0, 2])) < 1e-12 * Delta
    assert float((X.T @ eta + eta.T @ X).abs().max()) < 1e-12 and int(info[0, 0]) >= 1
    sopt = rb.options.default_option()
    sopt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=2, inner_maxiter=5, tolresid=0, maxtime=1e9)
    Y1 = torch.ones((n, p), dtype=torch.float64, device=dev)
    Xs, Ys, sm, _ = ss.solve(X, Y1, sopt)
    sm = sm.cpu().numpy()
    SM = rb._lib.SM
    assert np.isfinite(sm).all() and sm[0, SM["outer_iters"]] == 2 and sm[0, SM["inner_iters"]] >= 2
    assert float((Xs.T @ Xs - torch.eye(p, dtype=torch.float64, device=dev)).abs().max()) < 1e-13
    assert float(Xs.min()) > -eps and float(Ys.min()) > 0.0
    cost = lambda M: -float((M * (Z @ M)).sum())
    assert abs(sm[0, SM["cost"]] - cost(Xs)) < 1e-10 * abs(cost(Xs))
    ss.close()
