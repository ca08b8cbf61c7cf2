"""Structured description of a RIPTRM problem.

The reference hands the solver a `utils.NonlinearProblem` whose cost and constraints are opaque
autograd closures (src/solver/utils.py:33-77).  CUDA kernels need the structure behind them
(family + arrays).  It is obtained, in this order, from
  1. `problem.riptrm_structure` (set by a coordinator),
  2. the free variables of the reference coordinators' own closures (Z; alpha; X, XP, h; idx; row,
     col, ls/rs/c/sk), cross-checked numerically against the closures at the initial point,
and anything else raises: there is no CPU fallback for unknown problems.
"""
import types
from dataclasses import dataclass, field

import numpy as np

from . import _lib


@dataclass
class NonnegPCAStructure:
    """min -x'Zx on Sphere(n), x_i + eps >= 0 (src/NonnegPCA/coordinator.py:37-95)."""
    Z: np.ndarray
    x0: np.ndarray
    y0: np.ndarray
    eps: float = 0.0
    family: int = field(default=_lib.FAMILY_NONNEGPCA_SPHERE, init=False)

    @property
    def shape(self):
        n = self.Z.shape[0]
        return (n, 1, n)  # n, p, m

    def pack_x(self, x):
        return np.ascontiguousarray(x, dtype=np.float64).reshape(-1)

    def unpack_x(self, flat):
        return np.array(flat, dtype=np.float64)

    @property
    def typical_dist(self):
        return np.pi


@dataclass
class NonnegPCAStiefelStructure:
    """min -tr(X'ZX) on Stiefel(n, p), X_ij + eps >= 0 -- BASELINE config 4 as written (no reference coordinator: the
    reference's NonnegPCA lives on the sphere).  One run with an n x p matrix iterate; `y0` holds the
    n*p multipliers in row-major constraint order (flat or n x p)."""
    Z: np.ndarray
    x0: np.ndarray
    y0: np.ndarray
    eps: float = 0.01
    family: int = field(default=_lib.FAMILY_NONNEGPCA_STIEFEL, init=False)

    @property
    def shape(self):
        n, p = self.x0.shape
        return (n, p, n * p)

    @property
    def typical_dist(self):
        return np.sqrt(self.x0.shape[1])


@dataclass
class RosenbrockStructure:
    """Quadratic chain on Grassmann(n,k) (src/Rosenbrock/coordinator.py:33-91)."""
    n: int
    k: int
    alpha: float
    x0: np.ndarray
    y0: np.ndarray
    offset: float = 0.01
    family: int = field(default=_lib.FAMILY_ROSENBROCK_GRASSMANN, init=False)

    @property
    def shape(self):
        return (self.n, self.k, self.n * self.k)

    def pack_x(self, x):
        return np.ascontiguousarray(x, dtype=np.float64).reshape(-1)

    def unpack_x(self, flat):
        return np.array(flat, dtype=np.float64).reshape(self.n, self.k)

    @property
    def typical_dist(self):
        return np.sqrt(self.k)


@dataclass
class StableIdStructure:
    """A=(J-R)Q on Product[Skew(d),SPD(d),SPD(d)] (src/StableIdentification/coordinator.py:34-179).
    `conspec` rows: [kind, row, col, a, b] with kind 0: -A+a, 1: A-a, 2: -(A-a)^2+b."""
    X: np.ndarray
    XP: np.ndarray
    h: float
    conspec: np.ndarray
    x0: list
    y0: np.ndarray
    family: int = field(default=_lib.FAMILY_STABLEID_PRODUCT, init=False)

    @staticmethod
    def conspec_from_constset(constset):
        """Expands dataset constset rows the way coordinator.py:132-152 does."""
        rows = []
        for r in np.atleast_2d(np.asarray(constset, dtype=float)):
            t = r[0]
            if t == 0 or t == 1:
                rows.append([0.0, r[1], r[2], r[3], 0.0])
                rows.append([1.0, r[1], r[2], r[4], 0.0])
            elif t == 2:
                rows.append([2.0, r[1], r[2], r[3], r[4] ** 2])
            else:
                raise ValueError("Invalid constraint type")
        return np.array(rows, dtype=np.float64)

    @property
    def shape(self):
        return (self.X.shape[0], 3, len(self.conspec))

    def pack_x(self, x):
        return np.concatenate([np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in x])

    def unpack_x(self, flat):
        d = self.X.shape[0]
        f = np.array(flat, dtype=np.float64)
        return [f[i * d * d:(i + 1) * d * d].reshape(d, d) for i in range(3)]

    @property
    def typical_dist(self):
        d = self.X.shape[0]
        return np.sqrt(d * (d - 1) / 2 + d * (d + 1))


# ---------------------------------------------------------------------------------------------
# recognition of the reference coordinators' closures
# ---------------------------------------------------------------------------------------------
def _closure_vars(obj, max_nodes=200):
    """{free variable name: value} of the Python function(s) reachable from a (possibly wrapped)
    pymanopt Function object."""
    out, seen, stack = {}, set(), [obj]
    while stack and len(seen) < max_nodes:
        o = stack.pop()
        if id(o) in seen:
            continue
        seen.add(id(o))
        if isinstance(o, types.FunctionType) and o.__closure__:
            for name, cell in zip(o.__code__.co_freevars, o.__closure__):
                try:
                    v = cell.cell_contents
                except ValueError:
                    continue
                out.setdefault(name, v)
                if isinstance(v, types.FunctionType) or hasattr(v, "__wrapped__"):
                    stack.append(v)
        w = getattr(o, "__wrapped__", None)
        if w is not None:
            stack.append(w)
        d = getattr(o, "__dict__", None)
        if isinstance(d, dict) and not isinstance(o, (types.ModuleType, type)):
            for v in d.values():
                if isinstance(v, types.FunctionType) or hasattr(v, "__wrapped__") or (
                        callable(v) and hasattr(v, "__dict__") and not isinstance(v, type)):
                    stack.append(v)
    return out


def _float_consts(obj):
    consts, seen, stack = [], set(), [obj]
    while stack and len(seen) < 200:
        o = stack.pop()
        if id(o) in seen:
            continue
        seen.add(id(o))
        if isinstance(o, types.FunctionType):
            consts += [c for c in o.__code__.co_consts if isinstance(c, float)]
            for cell in (o.__closure__ or ()):
                try:
                    stack.append(cell.cell_contents)
                except ValueError:
                    pass
        w = getattr(o, "__wrapped__", None)
        if w is not None:
            stack.append(w)
        d = getattr(o, "__dict__", None)
        if isinstance(d, dict) and not isinstance(o, (types.ModuleType, type)):
            stack += [v for v in d.values() if callable(v) and not isinstance(v, type)]
    return consts


def _original_constraints(problem):
    cons = getattr(problem, "_original_ineqconstraints", None)
    if cons is None:
        cons = problem.ineqconstraints_all
    return list(cons)


def structure_from_problem(problem):
    """Returns the structure of a reference-style NonlinearProblem or raises NotImplementedError."""
    if isinstance(problem, (NonnegPCAStructure, NonnegPCAStiefelStructure, RosenbrockStructure, StableIdStructure)):
        return problem
    st = getattr(problem, "riptrm_structure", None)
    if st is not None:
        return st
    man = problem.manifold
    mname = type(man).__name__
    cost = getattr(problem, "_original_cost", None) or problem.cost
    cv = _closure_vars(cost)
    cons = _original_constraints(problem)
    x0 = problem.initialpoint
    y0 = np.asarray(problem.initialineqLagmult, dtype=np.float64)
    if mname == "Sphere" and "Z" in cv:
        Z = np.asarray(cv["Z"], dtype=np.float64)
        n = Z.shape[0]
        idx = [_closure_vars(c).get("idx") for c in cons]
        if len(cons) != n or idx != list(range(n)):
            raise NotImplementedError("Sphere problem whose constraints are not g_i = -x_i")
        st = NonnegPCAStructure(Z=Z, x0=np.asarray(x0, dtype=np.float64), y0=y0)
        _verify(problem, st, lambda x: -x @ Z @ x, lambda i, x: -x[i])
        return st
    if mname == "Grassmann" and "alpha" in cv:
        n, k = np.asarray(x0).shape
        idx = [_closure_vars(c).get("idx") for c in cons]
        if idx != list(range(n * k)):
            raise NotImplementedError("Grassmann problem whose constraints are not -vec(X)_i - offset")
        offs = [c for c in _float_consts(cons[0]) if c > 0]
        offset = offs[0] if offs else 0.01
        st = RosenbrockStructure(n=n, k=k, alpha=float(cv["alpha"]), x0=np.asarray(x0, dtype=np.float64), y0=y0,
                                 offset=float(offset))
        _verify(problem, st, None, lambda i, x: -x.flatten()[i] - offset)
        return st
    if mname == "Product" and all(k in cv for k in ("X", "XP", "h")):
        rows = []
        for c in cons:
            v = _closure_vars(c)
            code = getattr(getattr(c, "_function", None) or getattr(c, "__wrapped__", None) or c, "__name__", "")
            if "sk" in v and "c" in v:
                rows.append([2.0, v["row"], v["col"], v["c"], v["sk"]])
            elif "ls" in v and "rs" in v:
                # both one-box closures share (ls, rs) cells? they are separate closures: decide by name
                rows.append([0.0 if "ls" in code else 1.0, v["row"], v["col"], v["ls"] if "ls" in code else v["rs"], 0.0])
            elif "ls" in v:
                rows.append([0.0, v["row"], v["col"], v["ls"], 0.0])
            elif "rs" in v:
                rows.append([1.0, v["row"], v["col"], v["rs"], 0.0])
            else:
                raise NotImplementedError("unrecognised StableIdentification constraint closure")
        st = StableIdStructure(X=np.asarray(cv["X"], dtype=np.float64), XP=np.asarray(cv["XP"], dtype=np.float64),
                               h=float(cv["h"]), conspec=np.array(rows, dtype=np.float64),
                               x0=[np.asarray(a, dtype=np.float64) for a in x0], y0=y0)

        def g(i, x):
            A = (x[0] - x[1]) @ x[2]
            kind, r, c, a, b = st.conspec[i]
            v = A[int(r), int(c)]
            return -v + a if kind == 0 else (v - a if kind == 1 else -(v - a) ** 2 + b)
        _verify(problem, st, None, g)
        return st
    raise NotImplementedError(
        f"riptrm_b200 has no kernel family for this problem (manifold {mname}); attach "
        "`problem.riptrm_structure` (see INTEGRATION.md).  There is no CPU fallback.")


def _verify(problem, st, costfun, consfun):
    """Cross-checks the recovered structure against the opaque closures at the initial point."""
    x0 = problem.initialpoint
    if costfun is not None:
        a, b = float(problem.cost(x0)), float(costfun(np.asarray(x0)))
        if abs(a - b) > 1e-10 * max(1.0, abs(a)):
            raise NotImplementedError("recovered cost structure does not reproduce problem.cost(x0)")
    cons = problem.ineqconstraints_all
    for i in (0, len(cons) // 2, len(cons) - 1):
        a, b = float(cons[i](x0)), float(consfun(i, x0))
        if abs(a - b) > 1e-10 * max(1.0, abs(a)):
            raise NotImplementedError("recovered constraint structure does not reproduce the closures at x0")
