"""The fast lane of the last launch gets SMs of its own WHILE IT RUNS, and that is asserted on hardware.  Both
forms of the lane -- a second kernel on a priority stream (default) and CTAs elected inside sphere_tmem2_kernel
(RIPTRM_FAST_LANE_IN_KERNEL=1: placement by construction) -- write a placement record per CTA: SM id, role, %globaltimer at
entry and exit (`riptrm_lane_placement`).  Results are bit-identical with the lane switched off."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solve(rb, Z, x0, y0, opt):
    bs = rb.BatchSolver.nonnegpca_from_arrays(Z, x0, y0)
    bs.set_options(opt, 0, 0)
    x, y, sm, _ = bs.solve()
    rec = (C.c_int * 1024)()
    tm = (C.c_ulonglong * 2048)()
    n = bs.lib.riptrm_lane_placement(bs.handle.h, rec, tm, 1024)
    ms = bs.kernel_ms
    bs.close()
    ctas = [(rec[i] >> 4, rec[i] & 15, tm[2 * i], tm[2 * i + 1]) for i in range(max(n, 0)) if rec[i] != 0]
    return x, y, sm, ctas, ms


def _overlaps_on_lane_sms(ctas):
    """(lane SMs, main CTAs that ran on a lane SM at the same time as its lane CTA)"""
    lanes = [c for c in ctas if c[1] == 2]
    clash = []
    for sm_id, _, t0, t1 in lanes:
        for c in ctas:
            if c[1] == 1 and c[0] == sm_id and c[2] < t1 and c[3] > t0:
                clash.append((sm_id, c))
    return sorted(c[0] for c in lanes), clash


@pytest.fixture(scope="module")
def workload():
    import riptrm_b200 as rb
    Z, x0, y0 = rb.datagen.nonnegpca_sweep(7000, 1024, 4)      # 4096 pairs: two warps per copy of S, lane in the last launch
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=30, inner_maxiter=1000, tolresid=0, maxtime=1e9)
    return rb, Z, x0, y0, opt


def test_lane_kernel_runs_on_sms_of_its_own(workload, monkeypatch):
    """Default form: the lane kernel's 8 CTAs (high-priority stream, launched first, 150 KB of shared memory each) overlap in
    time with the main kernel's 296 CTAs, on 8 distinct SMs, and no main CTA runs on such an SM while its lane CTA does."""
    rb, Z, x0, y0, opt = workload
    x, y, sm, ctas, ms = _solve(rb, Z, x0, y0, opt)
    lane_sms, clash = _overlaps_on_lane_sms(ctas)
    assert len(lane_sms) == 8 and len(set(lane_sms)) == 8, lane_sms
    assert sum(1 for c in ctas if c[1] == 1) == 296
    assert clash == [], clash
    lane_t0 = min(c[2] for c in ctas if c[1] == 2)
    lane_t1 = max(c[3] for c in ctas if c[1] == 2)
    main_t0 = min(c[2] for c in ctas if c[1] == 1)
    main_t1 = max(c[3] for c in ctas if c[1] == 1)
    assert lane_t0 < main_t1 and main_t0 < lane_t1                               # the two kernels did run concurrently
    monkeypatch.setenv("RIPTRM_SPHERE_NO_FAST_LANE", "1")
    x2, y2, sm2, _, ms2 = _solve(rb, Z, x0, y0, opt)
    keep = [i for i in range(16) if i != rb._lib.SM["trace_rows"]]
    assert np.array_equal(x, x2) and np.array_equal(y, y2) and np.array_equal(sm[:, keep], sm2[:, keep])
    assert ms < ms2, (ms, ms2)                                                   # and it pays: 36 vs 43 ms on this batch
    print(f"two-kernel lane {ms:.2f} ms, lane off {ms2:.2f} ms")


def test_in_kernel_lane_owns_its_sms_by_construction(workload, monkeypatch):
    """RIPTRM_FAST_LANE_IN_KERNEL=1: one launch; the first CTA on each of 8 SMs becomes a lane CTA, its neighbour steps aside,
    every other SM runs two main CTAs; same results bit for bit."""
    rb, Z, x0, y0, opt = workload
    ref = _solve(rb, Z, x0, y0, opt)
    monkeypatch.setenv("RIPTRM_FAST_LANE_IN_KERNEL", "1")
    x, y, sm, ctas, ms = _solve(rb, Z, x0, y0, opt)
    roles = {}
    for c in ctas:
        roles.setdefault(c[0], []).append(c[1])
    lane_sms = [s for s, r in roles.items() if 2 in r]
    assert len(lane_sms) == 8, roles
    for s in lane_sms:
        assert sorted(roles[s]) == [2, 3], roles[s]                              # the lane CTA and the one that stepped aside
    for s, r in roles.items():
        if s not in lane_sms:
            assert r == [1, 1], (s, r)
    keep = [i for i in range(16) if i != rb._lib.SM["trace_rows"]]
    assert np.array_equal(x, ref[0]) and np.array_equal(y, ref[1]) and np.array_equal(sm[:, keep], ref[2][:, keep])
    print(f"in-kernel lane {ms:.2f} ms, two-kernel lane {ref[4]:.2f} ms")
