"""Import alias: `import riptrm_b200` loads the package that lives in the directory
`riemannian-interior-point-trust-region-method_b200/` (a name Python cannot import directly)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "riemannian-interior-point-trust-region-method_b200")
_spec = importlib.util.spec_from_file_location(
    "riptrm_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["riptrm_b200"] = _mod
_spec.loader.exec_module(_mod)
