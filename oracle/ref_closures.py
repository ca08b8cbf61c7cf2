"""TEST INFRASTRUCTURE -- builds the reference's own `NonlinearProblem` (unmodified
coordinators on the oracle/shims stand-ins) so closed-form derivatives in
oracle/problems.py can be checked against exact AD of the reference's closures.
Build container only (/root/reference)."""
import os
import sys
import tempfile

from .run_reference import REFERENCE, REPO, load_cfg


def reference_problem(problem_name, overrides=None):
    scratch = tempfile.mkdtemp(prefix="riptrm_refp_")
    os.symlink(f"{REFERENCE}/src", f"{scratch}/src")
    os.symlink(f"{REFERENCE}/dataset", f"{scratch}/dataset")
    old = os.getcwd()
    os.chdir(scratch)
    sys.path[:0] = [f"{REPO}/oracle/shims", REPO, f"./src/{problem_name}", "./src/solver", "./src/base"]
    try:
        for m in ("coordinator", "simulator"):
            sys.modules.pop(m, None)
        import coordinator
        cfg = load_cfg(problem_name, overrides)
        return coordinator.Coordinator(cfg).run()
    finally:
        os.chdir(old)
