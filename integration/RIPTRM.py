"""Drop-in module for the reference's solver lookup.

`Simulator.set_solver(name)` does `importlib.import_module(name)` and `getattr(module, name)(option)`
(src/base/base_simulator.py:64-66) with './src/solver' APPENDED to sys.path (:5).  Put this directory ahead
of it (PYTHONPATH=<riptrm_b200 repo>/integration) and `solver_name: ["RIPTRM"]` resolves to the B200 path;
nothing in the reference changes.  `TRS_solver: 'tCG'`, `second_order_stationarity: False` as in the
reference's config_simulation.yaml files; anything else raises (no CPU fallback)."""
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from riptrm_b200 import Output, RIPTRM  # noqa: E402,F401
