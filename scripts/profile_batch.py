"""Small driver for ncu / timing runs of the batched NonnegPCA whole-solve kernel.

    python scripts/profile_batch.py [pairs=4096] [reps=3] [maxiter=30]
"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np

import riptrm_b200 as rb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 30
Z, x0, y0 = rb.datagen.nonnegpca_batch(0, B, 50)
opt = rb.options.default_option()
opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=maxiter, tolresid=0, maxtime=1e9,
           inner_maxiter=1000)
bs = rb.BatchSolver.nonnegpca_from_arrays(Z, x0, y0)
bs.set_options(opt, 0, 0)
SM = rb._lib.SM
for rep in range(reps):
    t = time.time()
    x, y, sm, _ = bs.solve()
    wall = time.time() - t
    ms = bs.kernel_ms
    tcg, aux, inner = sm[:, SM["tcg_iters"]].sum(), sm[:, SM["aux_hessvecs"]].sum(), sm[:, SM["inner_iters"]].sum()
    print(f"rep {rep}: kernel {ms:.2f} ms wall {wall * 1e3:.1f} ms -> {B / (ms * 1e-3):.0f} pairs/s, "
          f"tcg it/s {tcg / (ms * 1e-3):.3e}, inner {inner / B:.1f}/pair, tcg {tcg / B:.1f}/pair, aux {aux / B:.1f}/pair; "
          f"resid max {sm[:, SM['residual']].max():.2e} stop {set(sm[:, SM['stop_reason']])}")
