"""Device-side sweep generation (csrc/datagen.cuh): bit-exact against its NumPy restatement
(tests/helpers.py), the reference generator's law (src/NonnegPCA/generator.py:9-65) as statistics, and a solve fed straight
from the generated device buffers."""
import numpy as np
import pytest

from helpers import device_generator_twin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import riptrm_b200
    return riptrm_b200


@pytest.mark.parametrize("n,first,instances,points", [(50, 0, 3, 1), (50, 4095, 2, 4), (37, 123456789012, 2, 2), (128, 7, 1, 3)])
def test_device_generator_is_bit_identical_to_its_restatement(rb, n, first, instances, points):
    Z, x0, y0 = rb.datagen.nonnegpca_sweep_device(first, instances, points, dim=n)
    Z, x0, y0 = Z.cpu().numpy(), x0.cpu().numpy(), y0.cpu().numpy()
    for i in range(instances):
        Zt, xt = device_generator_twin(first + i, n, points)
        assert np.array_equal(Z[i], Zt)
        assert np.array_equal(x0[i * points:(i + 1) * points], xt)
    assert np.all(y0 == 1.0)


def test_generator_law(rb):
    """2000 instances of n = 50: support size floor(0.7 n), spike sqrt(snr)/|S| on the support, off-diagonal noise
    N(0, 1/n), diagonal noise N(0, 4/n), Z not symmetric, unit-norm strictly positive starting points."""
    n, B, snr = 50, 2000, 0.5
    Z, x0, _ = rb.datagen.nonnegpca_sweep_device(10, B, 2, dim=n)
    Z, x0 = Z.cpu().numpy(), x0.cpu().numpy()
    k = int(np.floor(0.7 * n))
    off = ~np.eye(n, dtype=bool)
    spike = np.sqrt(snr) / k                         # per entry far below the noise (0.02 vs 0.14): test the mean level
    assert abs(Z[:, off].mean() - spike * (k * (k - 1)) / (n * (n - 1))) < 4 * (1 / np.sqrt(n)) / np.sqrt(B * n * (n - 1))
    assert abs(Z[:, off].std() - 1 / np.sqrt(n)) < 0.01 / np.sqrt(n)
    d = np.diagonal(Z, axis1=1, axis2=2)
    assert abs(d.std() - 2 / np.sqrt(n)) < 0.02 * 2 / np.sqrt(n)
    assert np.abs(Z - Z.transpose(0, 2, 1))[:, off].mean() > 0.05
    assert np.allclose(np.linalg.norm(x0, axis=1), 1.0, atol=1e-15) and x0.min() > 0.0
    # instances are independent streams: no two share a matrix, and the draw does not depend on the batch it is part of
    assert len({Z[i].tobytes() for i in range(50)}) == 50
    Z2, _, _ = rb.datagen.nonnegpca_sweep_device(15, 3, 1, dim=n)
    assert np.array_equal(Z2.cpu().numpy(), Z[5:8])


def test_solve_from_generated_device_buffers(rb):
    """Generated buffers go straight into the solver (no host round trip); same results as solving host copies."""
    import torch
    n, inst, ipp = 50, 64, 2
    Zd, x0d, y0d = rb.datagen.nonnegpca_sweep_device(1000, inst, ipp, dim=n)
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=30, inner_maxiter=1000, tolresid=0, maxtime=1e9)
    host = rb.BatchSolver.nonnegpca_from_arrays(Zd.cpu().numpy(), x0d.cpu().numpy(), y0d.cpu().numpy())
    host.set_options(opt, 0, 0)
    xh, yh, smh, _ = host.solve()
    host.set_nonnegpca(Zd, rb._lib.DEVICE)
    xd, yd = torch.empty_like(x0d), torch.empty_like(y0d)
    smd = torch.empty((inst * ipp, rb._lib.SUMMARY_FIELDS), dtype=torch.float64, device=Zd.device)
    host.solve_device(x0d, y0d, xd, yd, smd, None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    SM = rb._lib.SM
    assert np.array_equal(xd.cpu().numpy(), xh) and np.array_equal(smd.cpu().numpy()[:, :15], smh[:, :15])
    assert (smh[:, SM["residual"]] < 1e-9).all()
    host.close()
