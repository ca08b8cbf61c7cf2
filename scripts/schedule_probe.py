"""Cumulative work of every pair at each outer iteration (for offline simulation of the two-launch schedule):
python scripts/schedule_probe.py FIRST OUT.npz"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import riptrm_b200 as rb

first = int(sys.argv[1])
Z, x0, y0 = rb.datagen.nonnegpca_sweep(first, 4096, 4)
bs = rb.BatchSolver.nonnegpca_from_arrays(Z, x0, y0)
SM = rb._lib.SM
work = np.zeros((31, x0.shape[0]))
for k in range(1, 31):
    opt = rb.options.default_option()
    opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=k, inner_maxiter=1000, tolresid=0, maxtime=1e9, schedule_split=-1)
    bs.set_options(opt, 0, 0)
    x, y, sm, _ = bs.solve()
    work[k] = sm[:, SM["tcg_iters"]] + 2.0 * sm[:, SM["inner_iters"]]
    if k == 30:
        print("single launch ms", bs.kernel_ms)
for split in (4, 6, 8, 10):
    opt.update(maxiter=30, schedule_split=split)
    bs.set_options(opt, 0, 0)
    bs.solve(); bs.solve()
    print("split", split, "ms", bs.kernel_ms)
np.savez_compressed(sys.argv[2], work=work)
