"""Device time of the bench step for a given first instance (which rank's share is slow?): python scripts/range_probe.py FIRST [INSTANCES=4096] [IPP=4]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import riptrm_b200 as rb

first = int(sys.argv[1]); inst = int(sys.argv[2]) if len(sys.argv) > 2 else 4096; ipp = int(sys.argv[3]) if len(sys.argv) > 3 else 4
Z, x0, y0 = rb.datagen.nonnegpca_sweep(first, inst, ipp)
opt = rb.options.default_option()
opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=30, inner_maxiter=1000, tolresid=0, maxtime=1e9)
bs = rb.BatchSolver.nonnegpca_from_arrays(Z, x0, y0)
bs.set_options(opt, 0, 0)
for r in range(3):
    x, y, sm, _ = bs.solve()
    SM = rb._lib.SM
    t = sm[:, SM["tcg_iters"]]
    print(first, "kernel_ms %.2f" % bs.kernel_ms, "tcg/pair mean %.0f max %.0f p99 %.0f" % (t.mean(), t.max(), np.percentile(t, 99)),
          "inner max %.0f" % sm[:, SM["inner_iters"]].max())
