"""Drop-in boundary, build container only: the UNMODIFIED reference coordinators (src/<Problem>/coordinator.py,
run on the oracle/shims stand-ins for pymanopt / autograd / hydra) produce their own `NonlinearProblem`; the
structure the CUDA path needs is recovered from its closures and cross-checked numerically.  Skipped where
/root/reference does not exist (the GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs /root/reference")

SCRIPT = r'''
import sys, json
import numpy as np
sys.path.insert(0, {repo!r})
from oracle.ref_closures import reference_problem
import riptrm_b200 as rb
problem = reference_problem({name!r})
st = rb.structure_from_problem(problem)
out = {{"type": type(st).__name__, "shape": list(st.shape), "family": st.family}}
if {name!r} == "NonnegPCA":
    Z = np.loadtxt("/root/reference/dataset/NonnegPCA/1/Z.csv")
    out["Z_equal"] = bool(np.array_equal(st.Z, Z))
    out["x0_equal"] = bool(np.array_equal(st.x0, np.loadtxt("/root/reference/dataset/NonnegPCA/1/initx_a.csv")))
elif {name!r} == "Rosenbrock":
    out["alpha"] = st.alpha; out["offset"] = st.offset
else:
    out["conspec_rows"] = len(st.conspec); out["kinds"] = sorted(set(st.conspec[:, 0].tolist())); out["N"] = st.X.shape[1]
ld = rb.io.load_structure({name!r}, "/root/reference/dataset", 1, "a")
same = type(ld) is type(st) and np.array_equal(np.hstack([np.ravel(a) for a in ([ld.x0] if not isinstance(ld.x0, list) else ld.x0)]),
                                               np.hstack([np.ravel(a) for a in ([st.x0] if not isinstance(st.x0, list) else st.x0)]))
if {name!r} == "NonnegPCA":
    same = same and np.array_equal(ld.Z, st.Z)
elif {name!r} == "Rosenbrock":
    same = same and ld.alpha == st.alpha and ld.shape == st.shape
else:
    same = same and np.array_equal(ld.X, st.X) and np.array_equal(ld.XP, st.XP) and np.array_equal(ld.conspec, st.conspec) and ld.h == st.h
out["loader_agrees"] = bool(same)
print("RESULT " + json.dumps(out))
'''


def _run(name):
    r = subprocess.run([sys.executable, "-c", SCRIPT.format(repo=REPO, name=name)], capture_output=True, text=True,
                       cwd=REPO, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_nonnegpca_problem_is_recognised():
    out = _run("NonnegPCA")
    assert out["type"] == "NonnegPCAStructure" and out["shape"] == [50, 1, 50] and out["family"] == 1
    assert out["Z_equal"] and out["x0_equal"] and out["loader_agrees"]


def test_rosenbrock_problem_is_recognised():
    out = _run("Rosenbrock")
    assert out["type"] == "RosenbrockStructure" and out["shape"] == [5, 3, 15]
    assert out["alpha"] == 1e7 and out["offset"] == 0.01 and out["loader_agrees"]


def test_stableid_problem_is_recognised():
    out = _run("StableIdentification")
    assert out["type"] == "StableIdStructure" and out["shape"] == [5, 3, 16]
    assert out["conspec_rows"] == 16 and out["kinds"] == [0.0, 1.0, 2.0] and out["N"] == 95
    assert out["loader_agrees"]


def test_reference_simulator_reaches_the_c_abi_through_the_dropin_module():
    """The reference's own Simulator (unmodified, on the stand-ins) resolves solver_name=[RIPTRM] to
    integration/RIPTRM.py, builds the option dict, hands over its NonlinearProblem, and the call reaches the CUDA
    library -- which, in this GPU-less container, must fail loudly (RiptrmError from the C ABI), not fall back."""
    script = r'''
import sys
sys.path.insert(0, {repo!r})
from oracle.run_reference import run_reference
try:
    run_reference("NonnegPCA", {{"solver_option.common.maxiter": 3, "solver_option.common.tolresid": 0}},
                  solver_path={repo!r} + "/integration")
    print("RESULT solved")
except Exception as e:
    print("RESULT", type(e).__name__, str(e)[:200])
'''.format(repo=REPO)
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, cwd=REPO, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1]
    import torch
    if torch.cuda.is_available():
        assert line == "RESULT solved"
    else:
        assert line.startswith("RESULT RiptrmError") and "riptrm error -2" in line, line


def test_output_goes_through_the_reference_save_output_unchanged():
    """Wire format: an `Output` built by the host mirror from solver trace rows (here the
    rows come from the C oracle, which emits the device's trace layout) is written by the reference's own
    `Simulator.save_output` (src/base/base_simulator.py:75-95) and read back: log.csv has the reference's columns in the
    reference's order (+ tcg_iters), one row per logged iteration, and the analyzers' filter
    `inner_status == "converged"` selects one row per outer iteration."""
    script = r'''
import sys, os, json, copy, tempfile
sys.path.insert(0, {repo!r})
import numpy as np, pandas as pd
from oracle.run_reference import load_cfg, REFERENCE
import riptrm_b200 as rb
from oracle.c import binding as detc
scratch = tempfile.mkdtemp(prefix="riptrm_save_")
os.symlink(REFERENCE + "/src", scratch + "/src"); os.symlink(REFERENCE + "/dataset", scratch + "/dataset")
os.chdir(scratch)
sys.path[:0] = [{repo!r} + "/oracle/shims", "./src/NonnegPCA", "./src/solver", "./src/base"]
import simulator
cfg = load_cfg("NonnegPCA", {{"solver_name": ["RIPTRM"]}})
sim = simulator.Simulator(cfg)
os.makedirs(cfg.output_path, exist_ok=True)
Z = np.loadtxt("dataset/NonnegPCA/1/Z.csv"); x0 = np.loadtxt("dataset/NonnegPCA/1/initx_a.csv"); y0 = np.loadtxt("dataset/NonnegPCA/1/initineqLagmult.csv")
K = 6
x, y, sm, tr = detc.solve(Z, x0, y0, {{"maxiter": K, "tolresid": 0}}, trace_capacity=128)
option = rb.options.default_option(); option.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=K)
option["stoppingcriterion"] = "Max iteration count reached; maxiter=6 after 0.00 seconds"
out = rb.Output(name="RIPTRM_tCG", x=x, option=option, log=rb.trace_to_log(tr), ineqLagmult=y, eqLagmult=[])
sim.save_output(out.name, copy.deepcopy(out))
log = pd.read_csv(cfg.output_path + "/RIPTRM_tCG_log.csv")
xs = np.loadtxt(cfg.output_path + "/RIPTRM_tCG_x.csv")
res = {{"columns": list(log.columns), "rows": len(log), "conv": int((log["inner_status"] == "converged").sum()),
       "x_equal": bool(np.array_equal(xs, x)), "iter_last": int(log["iteration"].iloc[-1]),
       "files": sorted(os.listdir(cfg.output_path))}}
print("RESULT " + json.dumps(res))
'''.format(repo=REPO)
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, cwd=REPO, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    ref_cols = ["iteration", "time", "cost", "distance", "residual", "gradnorm", "complviolation", "dualviolation",
                "manviolation", "maxviolation", "meanviolation", "mu", "num_inner", "inner_status", "TR_radius", "dxtype",
                "normdx", "minxfeasi", "minyfeasi", "compl", "mineigvalHw", "ared/pred", "radius_update", "dual_clipping",
                "maxabsLagmult"]
    assert res["columns"][:len(ref_cols)] == ref_cols and res["columns"][len(ref_cols):] == ["tcg_iters"]
    assert res["conv"] == 6 and res["iter_last"] == 6 and res["x_equal"]
    assert {"RIPTRM_tCG_log.csv", "RIPTRM_tCG_x.csv", "RIPTRM_tCG_option.csv", "RIPTRM_tCG_ineqLagmult.csv",
            "RIPTRM_tCG_eqLagmult.csv", "RIPTRM_tCG_name.csv"} <= set(res["files"])
