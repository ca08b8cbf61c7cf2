#!/usr/bin/env python
"""Small ncu targets (B200_PROFILING.md: keep the profiled command short): one warm-up solve and one measured solve of

    sphere     the bench workload's pairs (default 16384: 4096 instances x 4 initial points), device-resident, tCG protocol
    stableid   the 2048-starting-point StableIdentification sweep of bench.py's secondary leg
    exact      256 NonnegPCA pairs with TRS_solver='Exact_RepMat' + second-order test

    python scripts/profile_target.py sphere [--pairs 16384]
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

PROTOCOL = {"TRS_solver": "tCG", "second_order_stationarity": False, "maxiter": 30, "inner_maxiter": 1000, "tolresid": 0,
            "maxtime": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("target", choices=["sphere", "stableid", "exact"])
    ap.add_argument("--pairs", type=int, default=0)
    args = ap.parse_args()
    import torch
    import riptrm_b200 as rb
    from riptrm_b200 import _lib
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    if args.target in ("sphere", "exact"):
        ipp = 4
        B = args.pairs or (16384 if args.target == "sphere" else 256)
        I = B // ipp
        Z, x0, y0 = (np.empty((I, 50, 50)), np.empty((B, 50)), np.empty((B, 50)))
        rb.datagen.nonnegpca_sweep(0, I, ipp, 50, out=(Z, x0, y0))
        solver = rb.BatchSolver.nonnegpca_from_arrays(Z, x0, y0, device=0)
        option = rb.options.default_option()
        option.update(PROTOCOL)
        if args.target == "exact":
            option.update(TRS_solver="Exact_RepMat", second_order_stationarity=True)
        solver.set_options(option, 0, 0)
        Zd, x0d, y0d = (torch.from_numpy(a).to(dev) for a in (Z, x0, y0))
        xd, yd = torch.empty_like(x0d), torch.empty_like(y0d)
        smd = torch.empty((B, _lib.SUMMARY_FIELDS), dtype=torch.float64, device=dev)
        solver.set_nonnegpca(Zd, _lib.DEVICE)
        stream = torch.cuda.current_stream().cuda_stream
        for rep in range(2):
            solver.solve_device(x0d, y0d, xd, yd, smd, None, stream)
            torch.cuda.synchronize()
            sm = smd.cpu().numpy()
            print(json.dumps({"target": args.target, "rep": rep, "pairs": B, "kernel_ms": solver.kernel_ms,
                              "tcg_iters": float(sm[:, _lib.SM["tcg_iters"]].sum()),
                              "inner_iters": float(sm[:, _lib.SM["inner_iters"]].sum()),
                              "aux_hessvecs": float(sm[:, _lib.SM["aux_hessvecs"]].sum()),
                              "max_residual": float(sm[:, _lib.SM["residual"]].max())}))
        solver.close()
    else:
        with open(os.path.join(REPO, "tests", "golden", "datasets.json")) as f:
            d = {k: np.array(v, dtype=float) for k, v in json.load(f)["StableIdentification/1"].items()}
        Xs = [d[f"noisyX_{k}"] for k in range(1, 6)]
        X, XP = np.hstack([x[:, :-1] for x in Xs]), np.hstack([x[:, 1:] for x in Xs])
        conspec = rb.StableIdStructure.conspec_from_constset(d["constset"])
        base = [[d[f"init{c}_{pt}"] for c in "JRQ"] for pt in "abcdefghijklmnopqrst"]
        pairs = args.pairs or 2048
        pts = rb.datagen.stableid_more_initial_points(base, conspec, pairs, seed=5)
        sts = [rb.StableIdStructure(X=X, XP=XP, h=0.02, conspec=conspec, x0=p, y0=d["initineqLagmult"]) for p in pts]
        option = rb.options.default_option()
        option.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=25, tolresid=0, maxtime=1e9)
        bs = rb.BatchSolver(sts, device=0)
        bs.set_options(option, 0, 0)
        for rep in range(2):
            x, y, sm, _ = bs.solve()
            print(json.dumps({"target": "stableid", "rep": rep, "pairs": pairs, "kernel_ms": bs.kernel_ms,
                              "tcg_iters": float(sm[:, _lib.SM["tcg_iters"]].sum()),
                              "inner_iters": float(sm[:, _lib.SM["inner_iters"]].sum()),
                              "aux_hessvecs": float(sm[:, _lib.SM["aux_hessvecs"]].sum()),
                              "residual_below_2e-9": float((sm[:, _lib.SM["residual"]] < 2e-9).mean())}))
        bs.close()


if __name__ == "__main__":
    main()
