#!/usr/bin/env python
"""Throughput of the bench workload with D solves in flight (one handle + one stream each, round robin).

    python scripts/pipeline_probe.py [--first-instance I] [--steps K] [--depths 1,2,3,4] [--no-lane]

A step is as long as its slowest launch's tail (the launches of the longest-first schedule drain one after the other); with
several steps in flight the drain of one step's launch is filled by the next step's CTAs.  Prints one JSON line per setting."""
import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import riptrm_b200 as rb  # noqa: E402
from riptrm_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--first-instance", type=int, default=0)
ap.add_argument("--instances", type=int, default=4096)
ap.add_argument("--initial-points", type=int, default=4)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--depths", default="1,2,3,4")
ap.add_argument("--lane", default="1,0", help="comma list: 1 = fast lane as shipped, 0 = RIPTRM_SPHERE_NO_FAST_LANE")
ap.add_argument("--splits", default="", help="RIPTRM_SPLITS values to try, ';'-separated (empty = default)")
args = ap.parse_args()

DIM = 50
I, ipp = args.instances, args.initial_points
B = I * ipp
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
rb.load_library()
Zh = np.empty((I, DIM, DIM))
x0h = np.empty((B, DIM))
y0h = np.empty((B, DIM))
rb.datagen.nonnegpca_sweep(args.first_instance, I, ipp, DIM, out=(Zh, x0h, y0h))
Zd, x0d, y0d = (torch.from_numpy(a).to(dev) for a in (Zh, x0h, y0h))
option = rb.options.default_option()
option.update({"TRS_solver": "tCG", "second_order_stationarity": False, "maxiter": 30, "inner_maxiter": 1000,
               "tolresid": 0, "maxtime": 1e9})
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SM = _lib.SM
ref_cost = None

for splits in args.splits.split(";"):
    if splits:
        os.environ["RIPTRM_SPLITS"] = splits
    else:
        os.environ.pop("RIPTRM_SPLITS", None)
    for lane in [int(v) for v in args.lane.split(",")]:
        if lane:
            os.environ.pop("RIPTRM_SPHERE_NO_FAST_LANE", None)
        else:
            os.environ["RIPTRM_SPHERE_NO_FAST_LANE"] = "1"
        for D in [int(v) for v in args.depths.split(",")]:
            solvers, streams, outs = [], [], []
            for d in range(D):
                s = rb.BatchSolver.nonnegpca_from_arrays(Zh[:1], x0h, y0h, device=0)   # Z is re-bound to the device copy below
                s.set_options(option, 0, 0)
                s.set_nonnegpca(Zd, _lib.DEVICE)
                solvers.append(s)
                streams.append(torch.cuda.Stream(dev))
                outs.append((torch.empty_like(x0d), torch.empty_like(y0d),
                             torch.empty((B, _lib.SUMMARY_FIELDS), dtype=torch.float64, device=dev)))

            def run(k):
                for s in range(k):
                    d = s % D
                    with torch.cuda.stream(streams[d]):
                        flush.fill_(s & 0xFF)
                        solvers[d].solve_device(x0d, y0d, outs[d][0], outs[d][1], outs[d][2], None, streams[d].cuda_stream)

            run(max(3, D))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for st in streams:
                st.wait_stream(torch.cuda.current_stream())
            run(args.steps)
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            cost = outs[0][2].cpu().numpy()[:, SM["cost"]]
            if ref_cost is None:
                ref_cost = cost
            same = all(np.array_equal(o[2].cpu().numpy()[:, SM["cost"]], ref_cost) for o in outs)
            print(json.dumps({"first_instance": args.first_instance, "splits": splits or "default", "lane": lane, "in_flight": D,
                              "ms_per_step": round(ms, 3), "pairs_per_s": round(B / (ms * 1e-3), 1), "bit_identical": bool(same)}),
                  flush=True)
            for s in solvers:
                s.close()
