"""TEST / BENCH INFRASTRUCTURE -- times the CPU oracle on a bounded sample of the bench workload
(bench.py's `cpu_baseline` leg and `--impl reference` arm; nothing on the product path imports this).

Two engines:
  "c"      oracle/c/riptrm_det.c (plain-C restatement with closed-form derivatives, one thread per
           pair via a process pool) when its shared object has been built (oracle/c/build.py);
  "numpy"  oracle/riptrm_oracle.py (per-constraint NumPy restatement that keeps the reference's
           evaluation structure).
Both follow src/solver/RIPTRM.py; the sample is pairs first_seed, first_seed+1, ... of the same
generator law the GPU arm uses.
"""
import multiprocessing as mp
import os
import time

import numpy as np


def _numpy_one(args):
    protocol, dim, seed = args
    from oracle.problems import NonnegPCAProblem, nonnegpca_generate_instance
    from oracle.riptrm_oracle import OracleRIPTRM
    Z, x0, y0 = nonnegpca_generate_instance(dim, seed=seed)
    opt = {k: v for k, v in protocol.items() if k not in ("TRS_solver", "second_order_stationarity", "maxtime")}
    opt["manviofun"] = NonnegPCAProblem.manviofun
    opt["save_inner_iteration"] = False
    o = OracleRIPTRM(opt)
    t = time.perf_counter()
    out = o.run(NonnegPCAProblem(Z, x0, y0))
    return time.perf_counter() - t, o.counters["tcg_hessvec"], float(out.log["residual"][-1])


def _c_available():
    try:
        from oracle.c import binding
        return binding.available()
    except Exception:
        return False


def run_sample(protocol, dim, first_seed=0, target_seconds=20.0, threads=None, max_pairs=4096, engine=None,
               points_per_instance=1):
    """Solves pairs first_seed.. on `threads` host threads for about `target_seconds` of wall time.
    Returns {"pairs", "seconds", "tcg_iters", "threads", "kind", "sample", "max_residual"}."""
    threads = threads or os.cpu_count() or 1
    if engine is None:
        engine = "c" if _c_available() else "numpy"
    if engine == "c":
        from oracle.c import binding
        return binding.run_sample(protocol, dim, first_seed, target_seconds, threads, max_pairs, points_per_instance)
    threads = min(threads, 64, max_pairs)
    # the NumPy oracle needs ~10-20 s per pair: one pair per worker
    per_pair_guess = 15.0
    rounds = max(1, int(target_seconds / per_pair_guess))
    n = min(max_pairs, threads * rounds)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    ctx = mp.get_context("fork")
    t = time.perf_counter()
    with ctx.Pool(threads) as pool:
        res = pool.map(_numpy_one, [(protocol, dim, first_seed + i) for i in range(n)], chunksize=1)
    secs = time.perf_counter() - t
    return {"pairs": n, "seconds": secs, "tcg_iters": int(sum(r[1] for r in res)), "threads": threads,
            "kind": "port", "engine": "numpy",
            "sample": f"{n} pairs (seeds {first_seed}..{first_seed + n - 1}) of the bench workload, NumPy oracle "
                      f"(per-constraint restatement), {threads} worker processes, full protocol",
            "max_residual": max(r[2] for r in res)}
