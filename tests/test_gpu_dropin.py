"""The drop-in boundary EXECUTED on the B200: caller side as the reference writes it
(tests/dropin_standins.py: coordinators' closures, `Simulator.set_solver` / `run` / `save_output`) -> module `RIPTRM` resolved to
integration/RIPTRM.py -> closure recognition -> C ABI -> CUDA kernels -> `Output` -> the CSV file set.  The Output must equal
the structure route (`run_batch(structures=...)`) bit for bit, and the golden run of the unmodified reference inside the
windows tests/test_gpu_parity_protocol.py asserts."""
import os
import warnings

import numpy as np
import pandas as pd
import pytest

from conftest import load_golden
from helpers import DISCRETE_COLUMNS, first_discrete_mismatch

pytestmark = pytest.mark.gpu

COMMON = {"maxtime": 1e9, "tolresid": 0, "verbosity": 0, "wandb_logging": False}
SPECIFIC = {"TRS_solver": "tCG", "second_order_stationarity": False}   # src/*/config_simulation.yaml solver_option.RIPTRM


def _same_output(a, b):
    xa = a.x if isinstance(a.x, list) else [a.x]
    xb = b.x if isinstance(b.x, list) else [b.x]
    assert all(np.array_equal(u, v) for u, v in zip(xa, xb))
    assert np.array_equal(a.ineqLagmult, b.ineqLagmult)
    for col in a.log:
        if col == "time":
            continue
        assert len(a.log[col]) == len(b.log[col])
        for u, v in zip(a.log[col], b.log[col]):
            assert u == v or (u != u and v != v), col


@pytest.mark.parametrize("name,maxiter,golden", [("NonnegPCA", 40, "nonnegpca_1_a_K40"), ("Rosenbrock", 6, "rosenbrock_K6"),
                                                 ("StableIdentification", 25, "stableid_1_a_K25")])
def test_simulator_route_runs_on_the_gpu_and_equals_the_structure_route(name, maxiter, golden, datasets, tmp_path):
    import dropin_standins as D
    import riptrm_b200 as rb
    sim = D.SimulatorStandIn(name, datasets, str(tmp_path), dict(COMMON, maxiter=maxiter), SPECIFIC)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out, problem = sim.run()
    assert type(out).__name__ == "Output" and out.name == "RIPTRM_tCG"
    assert out.option["stoppingcriterion"].startswith(f"Max iteration count reached; maxiter={maxiter}")
    if name == "Rosenbrock":      # the simulator's callback adds two logging-only columns: reported, not dropped silently
        assert any("callbackfun" in str(x.message) for x in w)
        assert out.option["riptrm_b200_missing_log_columns"] == "second_order_residual,condition_number"
    else:
        assert not any("callbackfun" in str(x.message) for x in w)
    # the structure route on the same data: bit-identical Output
    st = rb.structure_from_problem(D.build_problem(name, datasets))
    ref = rb.RIPTRM(dict(COMMON, **SPECIFIC, maxiter=maxiter)).run_batch([None], structures=[st])[0]
    _same_output(out, ref)
    # the golden run of the unmodified reference: same row 0, same discrete trace inside the window
    g = load_golden(golden)
    G = dict(g["log"], tcg_iters=[None] + g["tcg_iters"])
    for col in ("cost", "residual", "gradnorm", "complviolation"):
        assert abs(out.log[col][0] - G[col][0]) <= 1e-12 * max(1.0, abs(G[col][0])), col
    first = first_discrete_mismatch(out.log, G, columns=DISCRETE_COLUMNS + ("tcg_iters",))
    assert first >= {"NonnegPCA": 32, "Rosenbrock": 40, "StableIdentification": 30}[name], first
    # the file set Simulator.save_output writes, read back
    files = set(os.listdir(tmp_path))
    assert {f"RIPTRM_tCG_{a}.csv" for a in ("name", "x", "option", "log", "ineqLagmult", "eqLagmult")} <= files
    log = pd.read_csv(tmp_path / "RIPTRM_tCG_log.csv")
    assert len(log) == len(out.log["iteration"]) and int((log["inner_status"] == "converged").sum()) >= min(maxiter, 6)
    assert list(log.columns)[:3] == ["iteration", "time", "cost"] and "tcg_iters" in log.columns
    if name == "NonnegPCA":
        assert np.array_equal(np.loadtxt(tmp_path / "RIPTRM_tCG_x.csv"), out.x)


def test_foreign_manviofun_is_rejected_on_the_simulator_route(datasets, tmp_path):
    import dropin_standins as D
    sim = D.SimulatorStandIn("NonnegPCA", datasets, str(tmp_path), dict(COMMON, maxiter=2), SPECIFIC)
    problem = D.build_problem("NonnegPCA", datasets)
    solver = sim.set_solver("RIPTRM")
    solver.option["manviofun"] = lambda problem, x: float(np.abs(x).sum())
    with pytest.raises(NotImplementedError, match="manviofun"):
        solver.run(problem)
