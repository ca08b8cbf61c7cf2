"""Timing / ncu driver for the COLUMNS family (config 4): lock-step tCG capped at a fixed iteration count.

    python scripts/profile_columns.py [n=20000] [p=10] [iters=40] [reps=3]
    RIPTRM_COLUMNS_STREAM_ONLY=1 python scripts/profile_columns.py ...   # bare S.V passes (diagnostic)
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch

import riptrm_b200 as rb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 10
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda", 0)
peak = (json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured") \
    if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else (6650.0, "fallback")
out = []
roof, extra = rb.columns_bench(n, p, dev, peak, out, tcg_iters=iters, reps=reps)
print(json.dumps({"stream_only": bool(os.environ.get("RIPTRM_COLUMNS_STREAM_ONLY")), "roofline": roof, "extra": extra}))
