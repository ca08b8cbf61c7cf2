import sys, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import riptrm_b200 as rb
from oracle.problems import nonnegpca_generate_instance
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t = time.time()
sts = []
for seed in range(B):
    Z, x0, y0 = nonnegpca_generate_instance(50, seed=seed)
    sts.append(rb.NonnegPCAStructure(Z=Z, x0=x0, y0=y0))
print("gen", time.time() - t)
opt = rb.options.default_option(); opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=30, tolresid=0, maxtime=1e9, inner_maxiter=1000)
bs = rb.BatchSolver(sts)
bs.set_options(opt, 0, 0)
for rep in range(3):
    t = time.time(); x, y, sm, _ = bs.solve(); wall = time.time() - t
    ms = bs.kernel_ms
    tcg = sm[:, rb._lib.SM["tcg_iters"]].sum(); aux = sm[:, rb._lib.SM["aux_hessvecs"]].sum(); inner = sm[:, rb._lib.SM["inner_iters"]].sum()
    print(f"rep {rep}: kernel {ms:.2f} ms wall {wall*1e3:.1f} ms -> {B/(ms*1e-3):.0f} inst/s, tcg it/s {tcg/(ms*1e-3):.3e}, inner {inner/B:.1f}/inst tcg {tcg/B:.1f}/inst; resid max {sm[:,1].max():.2e} stop {set(sm[:,14])}")
