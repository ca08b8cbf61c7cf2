"""The NumPy oracle against outputs of the UNMODIFIED reference (tests/golden/*.json, made by
tests/golden/make_golden.py) and the reference notebooks' known answers."""
import numpy as np
import pytest

from conftest import load_golden
from helpers import (DISCRETE_COLUMNS, FLOAT_COLUMNS, first_discrete_mismatch, max_rel_diff, nonnegpca_problem,
                     rosenbrock_problem, stableid_problem)
from oracle.problems import NonnegPCAProblem, RosenbrockProblem, StableIdentificationProblem
from oracle.riptrm_oracle import OracleRIPTRM


def _golden_log(name):
    g = load_golden(name)
    return g, dict(g["log"], tcg_iters=[None] + g["tcg_iters"])


def test_nonnegpca_trace_is_bit_identical_to_reference(datasets):
    """NonnegPCA instance 1 / init a (BASELINE config 1), first 14 outer iterations: EVERY column of the
    reference's log -- discrete and floating point -- plus the tCG iteration counts, bit for bit."""
    K = 14
    g, G = _golden_log("nonnegpca_1_a_K40")
    out = OracleRIPTRM({"maxiter": K, "tolresid": 0, "manviofun": NonnegPCAProblem.manviofun}).run(
        nonnegpca_problem(datasets))
    n = len(out.log["iteration"])
    assert n > 40 and G["iteration"][n - 1] == K and G["iteration"][n] == K + 1
    for col in DISCRETE_COLUMNS + ("tcg_iters",):
        assert out.log[col] == G[col][:n], col
    for col in FLOAT_COLUMNS:
        assert max_rel_diff(out.log, G, col, rows=n) == 0.0, col


def test_notebook_known_answers(datasets):
    """Iteration-0 rows stored in the reference's analyzer notebooks."""
    o = OracleRIPTRM({"maxiter": 0, "manviofun": NonnegPCAProblem.manviofun}).run(nonnegpca_problem(datasets))
    assert abs(o.log["residual"][0] - 4.986888432851818) < 1e-13      # NonnegPCA/analyzer.ipynb: 4.986888e+00
    o = OracleRIPTRM({"maxiter": 0, "manviofun": RosenbrockProblem.manviofun}).run(rosenbrock_problem())
    assert abs(o.log["cost"][0] / 5.000001e7 - 1) < 1e-6              # Rosenbrock/analyzer.ipynb row 0
    assert abs(o.log["residual"][0] / 2.0e7 - 1) < 1e-6
    assert abs(o.log["complviolation"][0] - 1.749714) < 1e-6


def test_rosenbrock_matches_reference_window():
    """alpha = 1e7 makes the trajectory chaotic in rounding (closed-form derivatives vs the reference's
    AD): identical discrete trace for the first 40 rows, cost to 1e-9 there."""
    g, G = _golden_log("rosenbrock_K6")
    out = OracleRIPTRM({"maxiter": 1, "inner_maxiter": 60, "tolresid": 0,
                        "manviofun": RosenbrockProblem.manviofun}).run(rosenbrock_problem())
    first = first_discrete_mismatch(out.log, G, columns=DISCRETE_COLUMNS + ("tcg_iters",))
    assert first >= 40, first
    assert max_rel_diff(out.log, G, "cost", rows=first) < 1e-9
    assert max_rel_diff(out.log, G, "TR_radius", rows=first) < 1e-9


@pytest.mark.parametrize("pt", ["a"])
def test_stableid_matches_reference_window(datasets, pt):
    g, G = _golden_log(f"stableid_1_{pt}_K25")
    out = OracleRIPTRM({"maxiter": 2, "tolresid": 0, "manviofun": StableIdentificationProblem.manviofun}).run(
        stableid_problem(datasets, pt))
    first = first_discrete_mismatch(out.log, G, columns=DISCRETE_COLUMNS + ("tcg_iters",))
    assert first >= 30, first
    assert max_rel_diff(out.log, G, "cost", rows=20) < 1e-8
