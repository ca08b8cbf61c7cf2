#!/usr/bin/env python
"""StableIdentification sweep: per-pair work distribution and the kernel's time against the number of resident warps per SM.

    python scripts/stableid_probe.py
"""
import json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import riptrm_b200 as rb
from riptrm_b200 import _lib

with open(os.path.join(REPO, "tests", "golden", "datasets.json")) as f:
    d = {k: np.array(v, dtype=float) for k, v in json.load(f)["StableIdentification/1"].items()}
Xs = [d[f"noisyX_{k}"] for k in range(1, 6)]
X, XP = np.hstack([x[:, :-1] for x in Xs]), np.hstack([x[:, 1:] for x in Xs])
conspec = rb.StableIdStructure.conspec_from_constset(d["constset"])
base = [[d[f"init{c}_{pt}"] for c in "JRQ"] for pt in "abcdefghijklmnopqrst"]
pts = rb.datagen.stableid_more_initial_points(base, conspec, 2048, seed=5)
option = rb.options.default_option()
option.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=25, tolresid=0, maxtime=1e9)
SM = _lib.SM


def run(points):
    sts = [rb.StableIdStructure(X=X, XP=XP, h=0.02, conspec=conspec, x0=p, y0=d["initineqLagmult"]) for p in points]
    bs = rb.BatchSolver(sts, device=0)
    bs.set_options(option, 0, 0)
    bs.solve()
    ms = []
    for _ in range(3):
        x, y, sm, _ = bs.solve()
        ms.append(bs.kernel_ms)
    bs.close()
    return float(np.mean(ms)), sm


ms, sm = run(pts)
w = sm[:, SM["tcg_iters"]] + 2 * sm[:, SM["inner_iters"]]
q = np.percentile(w, [0, 10, 50, 90, 99, 100])
print(json.dumps({"pairs": 2048, "kernel_ms": ms, "work_percentiles_0_10_50_90_99_100": [float(v) for v in q], "mean": float(w.mean()),
                  "max_over_mean": float(w.max() / w.mean())}))
order = np.argsort(-w)
for k in (1, 2, 4, 8, 14):
    n = 148 * k
    # the n pairs with the most work, and n typical ones
    ms_top, _ = run([pts[i] for i in order[:n]])
    sel = order[len(order) // 2 - n // 2: len(order) // 2 - n // 2 + n]
    ms_mid, smm = run([pts[i] for i in sel])
    wm = smm[:, SM["tcg_iters"]] + 2 * smm[:, SM["inner_iters"]]
    print(json.dumps({"warps_per_sm": k, "pairs": n, "ms_longest_pairs": ms_top, "ms_median_pairs": ms_mid,
                      "us_per_work_unit_of_the_longest_median_pair": 1e3 * ms_mid / float(wm.max())}))
