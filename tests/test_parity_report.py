"""Quantified parity of the kernel arithmetic with the reference's golden run (VERDICT r1 "what's weak" 1-2): the
per-outer-iteration table of profiles/parity_r02.md with its MEASURED bounds asserted.  CPU leg: the deterministic C
oracle, which the GPU tests hold bit-identical to the CUDA kernels (tier T1); tests/test_gpu_parity_protocol.py asserts
the same bounds on the GPU itself.

north_star asks for "identical outer and tCG iteration counts" and "objective, KKT residual, iterates to a relative
1e-8".  What holds, and why not more:
  * outer iteration counts: identical (the protocol fixes them);
  * trust-region iteration counts, stop reasons, radius updates, clipping flags: identical through outer iteration 19;
  * tCG iteration counts: identical through outer iteration 9, totals within 3 %;
  * objective at every outer iteration: 3e-10; final objective, iterate, multipliers: 1e-15;
  * KKT residual at the end of an outer iteration: 1e-8 while the tCG counts agree, 1e-6 through outer 19, then 5e-3:
    it is the residual of an INEXACT inner solve (tolerance mu), not a converged quantity;
and the reference's own arithmetic is no closer to its golden run once the rounding of its dot products changes
(exactly rounded instead of BLAS order): tCG counts leave at outer 11, the discrete trace at 20, the residual by 8e-7 /
1e-2 (`test_reference_arithmetic_is_equally_sensitive`).  The window is a property of the algorithm's conditioning.
"""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "scripts"))

import parity_report as pr  # noqa: E402

# measured (profiles/parity_r02.md) -> asserted
BOUNDS = {
    "discrete_window_outer": 19,      # identical inner status / tCG stop reason / radius update / clipping through here
    "tcg_window_outer": 9,            # identical tCG iteration counts through here
    "tcg_total_rel": 0.035,           # total tCG iterations of the 40 outer iterations (measured 2.96 %)
    "inner_total_abs": 2,             # total trust-region iterations (measured 342 vs 341)
    "cost_rel_all": 1e-9,             # converged objective, every outer iteration (measured 2.9e-10)
    "cost_rel_late": 1e-12,           # ... from outer iteration 16 on (measured 7.8e-14)
    "resid_rel_tcg_window": 1e-8,     # KKT residual at the end of outer iterations 1..9 (measured 9.5e-9)
    "resid_rel_discrete_window": 2e-6,  # ... 10..19 (measured 9.5e-7)
    "resid_rel_after": 6e-3,          # ... 20..39 (measured 4.8e-3)
    "resid_abs_floor": 2e-15,         # outer 40: both at the rounding floor of the residual (7.1e-15 vs 8.2e-15)
    "final_rel": 1e-8,                # north_star tolerance on the final objective / iterate / multipliers
}


def check_nonnegpca_bounds(stats, a, b):
    assert stats["first_discrete_mismatch_outer"] > BOUNDS["discrete_window_outer"], stats
    assert stats["first_tcg_mismatch_outer"] > BOUNDS["tcg_window_outer"], stats
    assert stats["outer_window_inner_counts"] >= BOUNDS["discrete_window_outer"], stats
    assert stats["outer_window_tcg_counts"] >= BOUNDS["tcg_window_outer"], stats
    assert len(a["outer"]) == len(b["outer"]) == 40 and a["status"] == b["status"] == ["converged"] * 40
    assert abs(stats["tcg_total_here"] - stats["tcg_total_ref"]) <= BOUNDS["tcg_total_rel"] * stats["tcg_total_ref"]
    assert abs(stats["inner_total_here"] - stats["inner_total_ref"]) <= BOUNDS["inner_total_abs"]
    w = BOUNDS["tcg_window_outer"]
    assert np.array_equal(a["radius"][:w], b["radius"][:w])          # radii are exact while the tCG counts agree
    relc = np.abs(a["cost"] - b["cost"]) / np.abs(b["cost"])
    assert relc.max() < BOUNDS["cost_rel_all"] and relc[15:].max() < BOUNDS["cost_rel_late"], relc
    relr = np.abs(a["residual"] - b["residual"]) / b["residual"]
    assert relr[:9].max() <= BOUNDS["resid_rel_tcg_window"], relr[:9]
    assert relr[9:19].max() <= BOUNDS["resid_rel_discrete_window"], relr[9:19]
    assert relr[19:39].max() <= BOUNDS["resid_rel_after"], relr[19:39]
    assert abs(a["residual"][39] - b["residual"][39]) <= BOUNDS["resid_abs_floor"]
    assert stats["x_maxabs_diff"] < BOUNDS["final_rel"] and stats["y_rel_diff"] < BOUNDS["final_rel"]
    assert stats["final_cost_rel_diff"] < BOUNDS["final_rel"]


def test_c_oracle_per_outer_bounds():
    lines, stats, (a, b) = pr.nonnegpca_table("c")
    check_nonnegpca_bounds(stats, a, b)


def test_reference_arithmetic_is_equally_sensitive():
    """The reference's own arithmetic (NumPy oracle: bit-identical to the golden run as committed) under a different
    rounding of its dot products leaves the golden run where the kernels do -- so the windows asserted above are what
    any faithful implementation, the reference on another BLAS included, can reproduce."""
    rows = {r["variant"]: r for r in pr.reference_rounding_sensitivity()}
    base = rows["numpy BLAS dot (the oracle as committed)"]
    assert base["first_discrete_mismatch_outer"] is None and base["first_tcg_mismatch_outer"] is None
    assert base["tcg_total"] == 4194 and base["final_x_maxabs_diff"] == 0.0
    for name in ("exactly rounded dot (math.fsum)", "BLAS dot, reversed operand order"):
        r = rows[name]
        assert 9 < r["first_tcg_mismatch_outer"] <= 14, r            # kernels: 10
        assert 16 < r["first_discrete_mismatch_outer"] <= 22, r      # kernels: 20
        assert r["max_residual_rel_diff_outer_1_19"] > 1e-8          # the 1e-8 residual target fails for the reference too
        assert r["max_residual_rel_diff_outer_20_39"] > 1e-3
        assert r["final_x_maxabs_diff"] < 1e-12


def test_committed_report_is_current():
    """profiles/parity_r02.md is the committed output of scripts/parity_report.py: its NonnegPCA summary block must be
    what the script computes now (regenerate with `python scripts/parity_report.py`)."""
    path = os.path.join(REPO, "profiles", "parity_r02.md")
    assert os.path.exists(path), "run scripts/parity_report.py"
    text = open(path).read()
    _, stats, _ = pr.nonnegpca_table("c")
    for key in ("first_discrete_mismatch_row", "first_tcg_mismatch_row", "tcg_total_here", "tcg_total_ref",
                "inner_total_here"):
        assert f'"{key}": {stats[key]}' in text, key
