"""TEST INFRASTRUCTURE -- runs the UNMODIFIED reference (/root/reference) in this
container on the stand-in third-party packages of oracle/shims/, through the
reference's own `Simulator(cfg).run()` entry (src/<Problem>/simulator.py), and
returns the solver `Output` plus the per-call tCG iteration counts.

`/root/reference` is read-only and its modules use paths relative to the repo
root (`./src/base`, `dataset/...`, `intermediate/...`), so the run happens in a
scratch directory holding symlinks `src` and `dataset` into the reference.

The only instrumentation is a counting wrapper around the module-level function
`RIPTRM.truncated_conjugate_gradient` (src/solver/RIPTRM.py:41), because
`compute_direction` discards the iteration index `j` (:450).

Never imported by the product; cannot run on the GPU box (no /root/reference).
"""
import copy
import os
import re
import sys
import tempfile

import yaml

REFERENCE = "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class AttrDict(dict):
    """dict with attribute access (what the reference expects from an omegaconf DictConfig)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _to_attr(o):
    if isinstance(o, dict):
        return AttrDict({k: _to_attr(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_to_attr(v) for v in o]
    return o


def _yaml_float_loader():
    # PyYAML reads `1e7`/`1e-16` as strings (YAML 1.1 wants a dot); omegaconf reads floats.
    loader = yaml.SafeLoader
    loader.add_implicit_resolver(
        "tag:yaml.org,2002:float",
        re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)[eE][-+]?\d+$"),
        list("-+0123456789."),
    )
    return loader


def load_cfg(problem_name, overrides=None):
    path = f"{REFERENCE}/src/{problem_name}/config_simulation.yaml"
    with open(path) as f:
        raw = yaml.load(f, Loader=_yaml_float_loader())
    raw.pop("hydra", None)
    cfg = _to_attr(raw)
    for key, val in (overrides or {}).items():
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            if p not in node:
                node[p] = AttrDict()
            node = node[p]
        node[parts[-1]] = _to_attr(val)

    def interp(s):
        return re.sub(r"\$\{(\w+)\}", lambda m: str(cfg[m.group(1)]), s)

    def walk(node):
        for k, v in list(node.items()):
            if isinstance(v, str):
                node[k] = interp(v)
            elif isinstance(v, dict):
                walk(v)
    walk(cfg)
    return cfg


def run_reference(problem_name, overrides=None, solver_name="RIPTRM", solver_path=None, extra_option=None):
    """Returns (output, tcg_iters:list[int], scratch_dir).  `solver_path`: a directory put ahead of the
    reference's ./src/solver on sys.path (tests/test_dropin_recognition.py uses it to let the reference's
    Simulator pick up integration/RIPTRM.py instead of its own solver module)."""
    scratch = tempfile.mkdtemp(prefix="riptrm_ref_")
    os.symlink(f"{REFERENCE}/src", f"{scratch}/src")
    os.symlink(f"{REFERENCE}/dataset", f"{scratch}/dataset")
    old_cwd = os.getcwd()
    os.chdir(scratch)
    sys.path[:0] = ([solver_path] if solver_path else []) + [
        f"{REPO}/oracle/shims", REPO, f"./src/{problem_name}", "./src/solver", "./src/base"]
    try:
        import simulator  # the reference's src/<problem>/simulator.py
        import RIPTRM as ref_riptrm  # the reference's src/solver/RIPTRM.py (or the drop-in under solver_path)

        tcg_iters = []
        if hasattr(ref_riptrm, "truncated_conjugate_gradient"):
            orig_tcg = ref_riptrm.truncated_conjugate_gradient

            def counting_tcg(*a, **k):
                eta, Heta, j, stop = orig_tcg(*a, **k)
                tcg_iters.append(int(j) + 1)
                return eta, Heta, j, stop

            ref_riptrm.truncated_conjugate_gradient = counting_tcg

        ov = {"solver_name": [solver_name]}
        ov.update(overrides or {})
        cfg = load_cfg(problem_name, ov)
        sim = simulator.Simulator(cfg)
        if extra_option:
            orig_add = sim.add_solver_option

            def add_solver_option(option):
                option = orig_add(option)
                option.update(extra_option)
                return option

            sim.add_solver_option = add_solver_option
        outputs = {}
        orig_save = sim.save_output

        def capture(name, output):
            outputs[name] = output
            orig_save(name, copy.deepcopy(output))

        sim.save_output = capture
        sim.run()
        (name, output), = outputs.items()
        return output, tcg_iters, scratch
    finally:
        os.chdir(old_cwd)
