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
    assert out["Z_equal"] and out["x0_equal"]


def test_rosenbrock_problem_is_recognised():
    out = _run("Rosenbrock")
    assert out["type"] == "RosenbrockStructure" and out["shape"] == [5, 3, 15]
    assert out["alpha"] == 1e7 and out["offset"] == 0.01


def test_stableid_problem_is_recognised():
    out = _run("StableIdentification")
    assert out["type"] == "StableIdStructure" and out["shape"] == [5, 3, 16]
    assert out["conspec_rows"] == 16 and out["kinds"] == [0.0, 1.0, 2.0] and out["N"] == 95


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
