"""TEST INFRASTRUCTURE for the GPU box, where /root/reference does not exist: the CALLER side of the drop-in boundary
restated -- what the reference's coordinators and `Simulator` do around `RIPTRM(option).run(problem)` -- so that the route
Simulator -> integration/RIPTRM.py -> closure recognition -> C ABI -> CUDA -> Output -> save_output executes on hardware.

  * `NonlinearProblemStandIn`: the attributes `utils.NonlinearProblem.__init__` sets (src/solver/utils.py:33-77) on top of
    the stand-in `pymanopt.Problem` of oracle/shims (pymanopt itself is absent from the image);
  * `build_problem(name, datasets, initialpoint)`: cost / constraint closures written the way the three coordinators write them
    (src/NonnegPCA/coordinator.py:37-95, src/Rosenbrock/coordinator.py:33-91, src/StableIdentification/coordinator.py:34-179),
    with the same free-variable names, decorated with `pymanopt.function.autograd(manifold)`, fed from
    tests/golden/datasets.json (the reference's dataset CSVs);
  * `SimulatorStandIn`: `set_solver` (importlib lookup of module `RIPTRM`, option = common | solver-specific |
    add_solver_option, src/base/base_simulator.py:51-67), the per-workload `add_solver_option` (manviofun / callbackfun of the
    three simulator.py files), `run` (solver.run(copy.deepcopy(problem)), src/NonnegPCA/simulator.py:38) and `save_output`
    (base_simulator.py:75-95).

In the build container tests/test_dropin_recognition.py runs the same route with the UNMODIFIED reference modules.
"""
import copy
import csv
import importlib
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(REPO, "oracle", "shims")
if SHIMS not in sys.path:
    sys.path.insert(0, SHIMS)

import pymanopt  # noqa: E402  (the stand-in package of oracle/shims)
import pymanopt.function  # noqa: E402,F401
import pymanopt.manifolds  # noqa: E402,F401


class NonlinearProblemStandIn(pymanopt.Problem):
    """utils.NonlinearProblem's constructor (src/solver/utils.py:33-77): the attributes a solver reads."""

    def __init__(self, manifold, cost, ineqconstraints=(), eqconstraints=(), initialpoint=None,
                 initialineqLagmult=np.array([]), initialeqLagmult=np.array([])):
        super().__init__(manifold, cost)
        self._original_ineqconstraints = list(ineqconstraints)
        self.num_ineqconstraints = len(self._original_ineqconstraints)
        self.has_ineqconstraints = self.num_ineqconstraints > 0
        self._ineqconstraints = [self._wrap_function(c) for c in self._original_ineqconstraints]
        self._original_eqconstraints = list(eqconstraints)
        self.num_eqconstraints = len(self._original_eqconstraints)
        self.has_eqconstraints = self.num_eqconstraints > 0
        self._eqconstraints = [self._wrap_function(c) for c in self._original_eqconstraints]
        self.initialpoint = initialpoint
        self.initialineqLagmult = initialineqLagmult
        self.initialeqLagmult = initialeqLagmult

    def ineqconstraints(self, index):
        return self._ineqconstraints[index]

    @property
    def ineqconstraints_all(self):
        return [self.ineqconstraints(i) for i in range(self.num_ineqconstraints)]


# ---------------------------------------------------------------------------------------------------------------------
# the three coordinators' closures
# ---------------------------------------------------------------------------------------------------------------------
def _nonnegpca(d, initialpoint):
    dim = int(d["Z"].shape[0])
    mani = pymanopt.manifolds.sphere.Sphere(dim) if hasattr(pymanopt.manifolds, "sphere") else pymanopt.manifolds.Sphere(dim)
    Z = d["Z"]

    @pymanopt.function.autograd(mani)
    def costfun(point):
        return - point @ Z @ point

    def indexdecorated_nonnegfun(idx):
        @pymanopt.function.autograd(mani)
        def nonnegfun(point):
            return -point[idx]
        return nonnegfun

    constraint = [indexdecorated_nonnegfun(idx) for idx in range(dim)]
    return NonlinearProblemStandIn(mani, costfun, constraint, [], d[f"initx_{initialpoint}"].copy(),
                                   d["initineqLagmult"].copy(), np.array([]))


def _rosenbrock(n=5, k=3, alpha=1e7):
    mani = pymanopt.manifolds.Grassmann(n, k)

    @pymanopt.function.autograd(mani)
    def matrixrosenbrockfun(point):
        vectorized = point.flatten()
        num = len(vectorized)
        val = 0
        for i in range(num - 1):
            val = val + alpha * (vectorized[i + 1] - vectorized[i]) ** 2 + (1 - vectorized[i]) ** 2
        return val

    def indexdecorated_nonnegfun(idx):
        @pymanopt.function.autograd(mani)
        def nonnegfun(point):
            vectorized = point.flatten()
            return -vectorized[idx] - 0.01
        return nonnegfun

    constraint = [indexdecorated_nonnegfun(idx) for idx in range(n * k)]
    return NonlinearProblemStandIn(mani, matrixrosenbrockfun, constraint, [], np.abs(np.eye(n)[:, :k]), np.ones(n * k),
                                   np.array([]))


def _stableid(d, initialpoint, h=0.02, Xset=(1, 2, 3, 4, 5)):
    dim = int(d["initJ_a"].shape[0])
    mani = pymanopt.manifolds.product.Product([pymanopt.manifolds.SkewSymmetric(dim),
                                               pymanopt.manifolds.SymmetricPositiveDefinite(dim),
                                               pymanopt.manifolds.SymmetricPositiveDefinite(dim)])
    X = np.hstack([d[f"noisyX_{k}"][:, :-1] for k in Xset])
    XP = np.hstack([d[f"noisyX_{k}"][:, 1:] for k in Xset])
    N = X.shape[1]

    @pymanopt.function.autograd(mani)
    def costfun(J, R, Q):
        A = (J - R) @ Q
        Atilde = np.eye(dim) + h * A
        XPminusAtildeX = XP - Atilde @ X
        val = np.trace(XPminusAtildeX @ XPminusAtildeX.T) / N
        return val

    def build_onebox_constfuns(row, col, ls, rs):
        row = int(row)
        col = int(col)

        @pymanopt.function.autograd(mani)
        def onebox_lsconstfun(J, R, Q):
            A = (J - R) @ Q
            return -A[row, col] + ls

        @pymanopt.function.autograd(mani)
        def onebox_rsconstfun(J, R, Q):
            A = (J - R) @ Q
            return A[row, col] - rs
        return onebox_lsconstfun, onebox_rsconstfun

    def build_twobox_constfun(row, col, c, k):
        row = int(row)
        col = int(col)
        sk = k ** 2

        @pymanopt.function.autograd(mani)
        def twobox_constfun(J, R, Q):
            A = (J - R) @ Q
            return -(A[row, col] - c) ** 2 + sk
        return twobox_constfun

    constset = d["constset"]
    constraint = []
    for idx in range(constset.shape[0]):
        type_ = constset[idx, 0]
        if type_ == 0 or type_ == 1:
            ls_fun, rs_fun = build_onebox_constfuns(constset[idx, 1], constset[idx, 2], constset[idx, 3], constset[idx, 4])
            constraint += [ls_fun, rs_fun]
        elif type_ == 2:
            constraint.append(build_twobox_constfun(constset[idx, 1], constset[idx, 2], constset[idx, 3], constset[idx, 4]))
        else:
            raise ValueError("Invalid constraint type")
    x0 = [d[f"init{c}_{initialpoint}"].copy() for c in "JRQ"]
    return NonlinearProblemStandIn(mani, costfun, constraint, [], x0, d["initineqLagmult"].copy(), np.array([]))


def build_problem(name, datasets, initialpoint="a"):
    if name == "NonnegPCA":
        return _nonnegpca(datasets["NonnegPCA/1"], initialpoint)
    if name == "Rosenbrock":
        return _rosenbrock()
    if name == "StableIdentification":
        return _stableid(datasets["StableIdentification/1"], initialpoint)
    raise KeyError(name)


# ---------------------------------------------------------------------------------------------------------------------
# the simulators' option hooks (src/<Problem>/simulator.py `add_solver_option`)
# ---------------------------------------------------------------------------------------------------------------------
def _manvio_nonnegpca(problem, x):
    return np.linalg.norm(x) - 1


def _manvio_rosenbrock(problem, x):
    manvio = 0
    if np.linalg.matrix_rank(x) != problem.manifold._p:
        manvio = np.inf
    return manvio


def _callback_rosenbrock(problem, x, ineqLagmult, eqLagmult, eval):
    # src/Rosenbrock/simulator.py:100-105 adds two logging-only columns (second-order residual on a RANDOM basis)
    eval["second_order_residual"] = 0.0
    eval["condition_number"] = None
    return eval


def _manvio_stableid(problem, x):
    J, R, Q = x
    manvio = np.linalg.norm(J + J.T) + np.linalg.norm(R - R.T) + np.linalg.norm(Q - Q.T)
    if not np.all(np.linalg.eigvalsh(R) > 0) or not np.all(np.linalg.eigvalsh(Q) > 0):
        manvio = np.inf
    return manvio


ADD_SOLVER_OPTION = {
    "NonnegPCA": {"manviofun": _manvio_nonnegpca},
    "Rosenbrock": {"manviofun": _manvio_rosenbrock, "callbackfun": _callback_rosenbrock},
    "StableIdentification": {"manviofun": _manvio_stableid},
}


class SimulatorStandIn:
    """base_simulator.Simulator (src/base/base_simulator.py) around one solver run."""

    def __init__(self, problem_name, datasets, output_path, common, specific, initialpoint="a"):
        self.problem_name, self.datasets, self.output_path = problem_name, datasets, output_path
        self.common, self.specific, self.initialpoint = common, specific, initialpoint

    def set_solver(self, solver_name):
        option = copy.deepcopy(dict(self.common))                      # :56
        option.update(dict(self.specific))                              # :57-59, common before specific
        option.update(ADD_SOLVER_OPTION[self.problem_name])            # :61 add_solver_option
        integration = os.path.join(REPO, "integration")
        if integration not in sys.path:
            sys.path.insert(0, integration)                            # ahead of ./src/solver
        module_solver = importlib.import_module(solver_name)            # :64
        return getattr(module_solver, solver_name)(option)              # :65-66

    def save_output(self, solver_name, output):                         # :75-95
        import pandas as pd
        os.makedirs(self.output_path, exist_ok=True)
        for attr, content in vars(output).items():
            csvpath = f"{self.output_path}/{solver_name}_{attr}.csv"
            if isinstance(content, (np.matrix, np.ndarray)):
                np.savetxt(csvpath, content)
            elif isinstance(content, dict):
                for key, value in content.items():
                    if not isinstance(value, list):
                        content[f"{key}"] = [value]
                pd.DataFrame(content).to_csv(csvpath, index=False)
            else:
                with open(csvpath, "w") as csvfile:
                    csv.writer(csvfile).writerows(content)

    def run(self):
        problem = build_problem(self.problem_name, self.datasets, self.initialpoint)
        solver = self.set_solver("RIPTRM")
        output = solver.run(copy.deepcopy(problem))                    # src/NonnegPCA/simulator.py:38
        self.save_output(output.name, copy.deepcopy(output))
        return output, problem
