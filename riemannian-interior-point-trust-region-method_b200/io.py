"""The reference's on-disk formats on both sides of the hot path.

Inputs -- `dataset/<problem>/<instance>/*.csv`, written by the reference's generators with `np.savetxt`
(src/base/dataset_generator.py:39-50; whitespace separated, '%.18e') and read by its coordinators with
`np.loadtxt` (src/NonnegPCA/coordinator.py:39-95, src/StableIdentification/coordinator.py:53-179):
`load_structure` builds the structured problem description straight from those files, which is the second route
the solver boundary allows besides closure recognition (structure.py) -- no pymanopt / autograd / hydra involved.

Outputs -- `<output_path>/<solver>_{name,x,option,log,ineqLagmult,eqLagmult}.csv`, written by
`Simulator.save_output` (src/base/base_simulator.py:75-95): `save_output` writes the same files in the same way
(ndarray -> np.savetxt; dict -> one-row / many-row CSV through pandas; anything else -> csv.writer.writerows), so the
reference's analyzer notebooks read a GPU run like one of their own.
"""
import csv
import os

import numpy as np

from .structure import NonnegPCAStructure, RosenbrockStructure, StableIdStructure


def save_dataset(path, **arrays):
    """np.savetxt each array to <path>/<name>.csv (dataset_generator.Generator.save)."""
    os.makedirs(path, exist_ok=True)
    for name, content in arrays.items():
        np.savetxt(os.path.join(path, f"{name}.csv"), np.asarray(content))


def load_nonnegpca(dataset_root, instance, initialpoint="a"):
    """dataset/NonnegPCA/<instance>: dim, Z, initx_<initialpoint>, initineqLagmult (coordinator.py:39-95)."""
    d = os.path.join(dataset_root, "NonnegPCA", str(instance))
    Z = np.loadtxt(os.path.join(d, "Z.csv"))
    dim = int(np.loadtxt(os.path.join(d, "dim.csv")))
    if Z.shape != (dim, dim):
        raise ValueError(f"{d}: Z is {Z.shape}, dim.csv says {dim}")
    return NonnegPCAStructure(Z=Z, x0=np.loadtxt(os.path.join(d, f"initx_{initialpoint}.csv")),
                              y0=np.loadtxt(os.path.join(d, "initineqLagmult.csv")))


def load_stableid(dataset_root, instance, initialpoint="a", Xset=(1, 2, 3, 4, 5), h=0.02, is_X_noisy=True):
    """dataset/StableIdentification/<instance>: trajectories (noisy)X_k (columns t and t+1 give X and X'),
    constset, init{J,R,Q}_<initialpoint>, initineqLagmult (coordinator.py:53-179; config_simulation.yaml:10-12)."""
    d = os.path.join(dataset_root, "StableIdentification", str(instance))
    Xs, XPs = [], []
    for k in Xset:
        Xo = np.loadtxt(os.path.join(d, f"{'noisyX' if is_X_noisy else 'X'}_{k}.csv"))
        Xs.append(Xo[:, :-1])
        XPs.append(Xo[:, 1:])
    conspec = StableIdStructure.conspec_from_constset(np.loadtxt(os.path.join(d, "constset.csv")))
    return StableIdStructure(X=np.hstack(Xs), XP=np.hstack(XPs), h=float(h), conspec=conspec,
                             x0=[np.loadtxt(os.path.join(d, f"init{c}_{initialpoint}.csv")) for c in "JRQ"],
                             y0=np.loadtxt(os.path.join(d, "initineqLagmult.csv")))


def rosenbrock(n=5, k=3, alpha=1e7):
    """The Rosenbrock workload reads nothing from disk: x0 = I[:, :k], y0 = 1 (src/Rosenbrock/coordinator.py:78-91)."""
    return RosenbrockStructure(n=n, k=k, alpha=float(alpha), x0=np.abs(np.eye(n)[:, :k]), y0=np.ones(n * k))


def load_structure(problem_name, dataset_root, instance=1, initialpoint="a", **kw):
    if problem_name == "NonnegPCA":
        return load_nonnegpca(dataset_root, instance, initialpoint)
    if problem_name == "StableIdentification":
        return load_stableid(dataset_root, instance, initialpoint, **kw)
    if problem_name == "Rosenbrock":
        return rosenbrock(**kw)
    raise NotImplementedError(f"no kernel family for problem {problem_name!r}")


def save_output(output, output_path, solver_name=None):
    """Writes what `Simulator.save_output(solver_name, output)` writes (base_simulator.py:75-95)."""
    import pandas as pd
    os.makedirs(output_path, exist_ok=True)
    solver_name = solver_name or output.name
    for attr, content in vars(output).items():
        csvpath = os.path.join(output_path, f"{solver_name}_{attr}.csv")
        if isinstance(content, (np.matrix, np.ndarray)):
            np.savetxt(csvpath, content)
        elif isinstance(content, dict):
            content = {key: (value if isinstance(value, list) else [value]) for key, value in content.items()}
            pd.DataFrame(content).to_csv(csvpath, index=False)
        else:
            with open(csvpath, "w") as csvfile:
                csv.writer(csvfile).writerows(content)
    return output_path
