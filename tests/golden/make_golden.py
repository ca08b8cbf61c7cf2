"""Generates tests/golden/*.json by running the UNMODIFIED reference (/root/reference)
through its own Simulator entry on the third-party stand-ins of oracle/shims/
(see oracle/run_reference.py).  Run from the repo root, in the build container only:

    python tests/golden/make_golden.py [case ...]

Each case runs in its own process (the reference's per-problem modules share the
module names `simulator` / `coordinator`).  Wall-clock stopping is disabled
(maxtime=1e9) and the run is capped by `maxiter` (SURVEY.md App. C protocol).
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))

COMMON = {"solver_option.common.tolresid": 0, "solver_option.common.maxtime": 1e9}
CASES = {
    # name: (problem, overrides)
    "nonnegpca_1_a_K40": ("NonnegPCA", {"solver_option.common.maxiter": 40}),
    "rosenbrock_K6": ("Rosenbrock", {"solver_option.common.maxiter": 6}),
    "stableid_1_a_K25": ("StableIdentification", {"solver_option.common.maxiter": 25, "problem_initialpoint": "a"}),
    "stableid_1_b_K25": ("StableIdentification", {"solver_option.common.maxiter": 25, "problem_initialpoint": "b"}),
    "stableid_1_t_K25": ("StableIdentification", {"solver_option.common.maxiter": 25, "problem_initialpoint": "t"}),
    # round 2: a converged Rosenbrock run (the reference reaches cost 4.0000009514e7 by outer iteration 14) ...
    "rosenbrock_K14": ("Rosenbrock", {"solver_option.common.maxiter": 14}),
    "rosenbrock_K20": ("Rosenbrock", {"solver_option.common.maxiter": 20, "solver_option.RIPTRM.inner_maxiter": 2000}),
}
# ... and all 20 initial points of StableIdentification instance 1 (dataset/StableIdentification/1/
# init{J,R,Q}_{a..t}.csv) at 30 outer iterations (mu = 3.9e-12: the identified quantity A = (J-R)Q is converged to
# ~1e-10 there; at 25 it still moves by 1e-7), stored reduced: row 0 and the rows that end an outer iteration
for _pt in "abcdefghijklmnopqrst":
    CASES[f"stableid_1_{_pt}_K30"] = ("StableIdentification", {"solver_option.common.maxiter": 30,
                                                                 "solver_option.RIPTRM.inner_maxiter": 1000,
                                                                 "problem_initialpoint": _pt})
# round 2: the reference's class-default trust-region solver (TRS_solver='Exact_RepMat', second_order_stationarity=True,
# RIPTRM.py:323-324) with a deterministic `basisfun` (riptrm_b200/basis.py; the reference's default basis is random)
EXACT = {"solver_option.RIPTRM.TRS_solver": "Exact_RepMat", "solver_option.RIPTRM.second_order_stationarity": True,
         "solver_option.RIPTRM.inner_maxiter": 1000}
CASES["nonnegpca_1_a_exact_K40"] = ("NonnegPCA", dict(EXACT, **{"solver_option.common.maxiter": 40}))
CASES["rosenbrock_exact_K14"] = ("Rosenbrock", dict(EXACT, **{"solver_option.common.maxiter": 14}))
# StableIdentification: the Hessian of the Lagrangian is singular along the directions that leave A = (J-R)Q unchanged
# (smallest eigenvalue ~ -1e-9), so the second-order test `mineig >= -mu` can only pass while mu > 1e-9: 14 outer iterations
# (mu = 5e-6); beyond ~22 the reference's inner loop never converges again and runs into inner_maxiter
for _pt in "ab":
    CASES[f"stableid_1_{_pt}_exact_K14"] = ("StableIdentification", dict(EXACT, **{"solver_option.common.maxiter": 14,
                                                                                   "problem_initialpoint": _pt}))
REDUCED = {"rosenbrock_K14", "rosenbrock_K20"} | {f"stableid_1_{_pt}_K30" for _pt in "abcdefghijklmnopqrst"}
# columns whose values depend on the reference's unseeded RNG (Rosenbrock callback,
# src/Rosenbrock/simulator.py:52-57) or on wall-clock
SKIP_COLUMNS = ("time", "second_order_residual", "condition_number")


def _jsonable(v):
    if v is None or isinstance(v, (str, bool)):
        return v
    if isinstance(v, (np.bool_,)):
        return bool(v)
    if isinstance(v, (int, np.integer)):
        return int(v)
    if isinstance(v, (float, np.floating)):
        return float(v)
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (list, tuple)):
        return [_jsonable(u) for u in v]
    raise TypeError(type(v))


def run_case(name):
    sys.path.insert(0, REPO)
    from oracle.run_reference import run_reference

    problem, ov = CASES[name]
    overrides = dict(COMMON)
    overrides.update(ov)
    extra = None
    if overrides.get("solver_option.RIPTRM.TRS_solver") == "Exact_RepMat":
        import riptrm_b200
        from riptrm_b200.basis import deterministic_basisfun
        extra = {"basisfun": deterministic_basisfun}
    out, tcg, _ = run_reference(problem, overrides, extra_option=extra)
    log = {k: _jsonable(v) for k, v in out.log.items() if k not in SKIP_COLUMNS}
    rows_total = len(log["iteration"])
    inner_per_outer = None
    if name in REDUCED:
        # per-outer-iteration records only: row 0 + the last row of every outer iteration (`converged`, or the row at
        # which an inner run hit its cap); the number of trust-region iterations and tCG iterations per outer iteration
        # are kept as counts
        it = log["iteration"]
        keep = [0] + [i for i in range(1, rows_total) if i + 1 == rows_total or it[i + 1] != it[i]]
        K = max(it)
        inner_per_outer = [sum(1 for i in range(1, rows_total) if it[i] == k) for k in range(1, K + 1)]
        tcg_per_outer = [sum(tcg[i - 1] for i in range(1, rows_total) if it[i] == k) for k in range(1, K + 1)]
        log = {k: [v[i] for i in keep] for k, v in log.items()}
        tcg = tcg_per_outer
    doc = {
        "case": name,
        "problem": problem,
        "overrides": {k: v for k, v in overrides.items()},
        "generator": "tests/golden/make_golden.py (unmodified reference on oracle/shims stand-ins)",
        "solver_name": out.name,
        "stoppingcriterion": out.option["stoppingcriterion"].split(" after ")[0],
        "x": _jsonable(out.x),
        "ineqLagmult": _jsonable(out.ineqLagmult),
        "tcg_iters": tcg,
        "rows_total": rows_total,
        "inner_per_outer": inner_per_outer,
        "reduced": name in REDUCED,
        "log": log,
    }
    with open(os.path.join(HERE, f"{name}.json"), "w") as f:
        json.dump(doc, f)
    print(name, "rows", len(log["iteration"]), "tcg hessvecs", sum(tcg))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
    else:
        names = sys.argv[1:] or list(CASES)
        procs = [(n, subprocess.Popen([sys.executable, __file__, "--one", n], cwd=REPO)) for n in names]
        for n, p in procs:
            if p.wait() != 0:
                raise SystemExit(f"case {n} failed")
