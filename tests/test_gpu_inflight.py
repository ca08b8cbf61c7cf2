"""Several solves in flight on one GPU (bench.py --in-flight): handles are independent, so solves enqueued on
different streams -- their launches alternate on the SMs, each filling the other's drain -- return exactly what one solve at a
time returns."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_solves_in_flight_on_two_handles_are_bit_identical_to_one_at_a_time():
    import torch
    import riptrm_b200 as rb
    from riptrm_b200 import _lib
    dev = torch.device("cuda", 0)
    I, ipp, n = 1200, 4, 50          # 4800 pairs: more than the GPU holds at once, so the multi-launch schedule and the lane run
    B = I * ipp
    Z, x0, y0 = np.empty((I, n, n)), np.empty((B, n)), np.empty((B, n))
    rb.datagen.nonnegpca_sweep(7000, I, ipp, n, out=(Z, x0, y0))
    option = rb.options.default_option()
    option.update({"TRS_solver": "tCG", "second_order_stationarity": False, "maxiter": 30, "inner_maxiter": 1000,
                   "tolresid": 0, "maxtime": 1e9})
    Zd, x0d, y0d = (torch.from_numpy(a).to(dev) for a in (Z, x0, y0))
    solvers, streams, outs = [], [], []
    for d in range(2):
        s = rb.BatchSolver.nonnegpca_from_arrays(Z[:1], x0, y0, device=0)
        s.set_options(option, 0, 0)
        s.set_nonnegpca(Zd, _lib.DEVICE)
        solvers.append(s)
        streams.append(torch.cuda.Stream(dev))
        outs.append([(torch.empty_like(x0d), torch.empty_like(y0d),
                      torch.empty((B, _lib.SUMMARY_FIELDS), dtype=torch.float64, device=dev)) for _ in range(2)])
    # reference: one solve, alone
    solvers[0].solve_device(x0d, y0d, *outs[0][0], None, streams[0].cuda_stream)
    torch.cuda.synchronize()
    ref = [t.clone() for t in outs[0][0]]
    for t in outs[0][0]:
        t.zero_()
    # four solves, two per handle, enqueued alternately without any synchronisation in between
    for k in range(2):
        for d in range(2):
            solvers[d].solve_device(x0d, y0d, *outs[d][k], None, streams[d].cuda_stream)
    torch.cuda.synchronize()
    keep = [i for i in range(_lib.SUMMARY_FIELDS)]
    for d in range(2):
        for k in range(2):
            x, y, sm = outs[d][k]
            assert torch.equal(x, ref[0]) and torch.equal(y, ref[1])
            assert torch.equal(sm[:, keep], ref[2][:, keep])
    assert float((ref[2][:, _lib.SM["residual"]] <= 1e-8).double().mean()) == 1.0
    for s in solvers:
        s.close()
