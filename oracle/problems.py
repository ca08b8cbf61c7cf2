"""TEST INFRASTRUCTURE -- CPU oracle, not product code.

Closed-form restatement of the three workloads the reference solves with RIPTRM
(autograd, which differentiates the reference's closures, is absent here; the
closed forms are checked against finite differences and against exact AD of the
reference's OWN closures in tests/test_oracle_problems.py and
tests/golden/make_golden.py).

Each class exposes the accessor surface of the reference's
`utils.NonlinearProblem` (src/solver/utils.py:33-203): `cost`,
`euclidean_gradient`, `euclidean_hessian`, `riemannian_gradient`,
`riemannian_hessian`, `ineqconstraints_all`, `ineqconstraints_*_all`,
`initialpoint`, `initialineqLagmult`, `preconditioner`.

Conventions (src/solver/RIPTRM.py:576,721): constraints g_i(x) <= 0.
"""
import numpy as np

from . import manifolds as M


class OracleProblem:
    """Accessor surface of utils.NonlinearProblem (src/solver/utils.py:79-173)."""

    has_eqconstraints = False
    num_eqconstraints = 0
    eqconstraints_all = []

    def __init__(self, manifold, initialpoint, initialineqLagmult):
        self.manifold = manifold
        self.initialpoint = initialpoint
        self.initialineqLagmult = np.asarray(initialineqLagmult, dtype=float)
        self.num_ineqconstraints = len(self.initialineqLagmult)
        self.has_ineqconstraints = self.num_ineqconstraints > 0

    # --- to be provided by subclasses ------------------------------------
    def cost(self, x):
        raise NotImplementedError

    def euclidean_gradient(self, x):
        raise NotImplementedError

    def euclidean_hessian(self, x, v):
        raise NotImplementedError

    def ineq(self, i, x):
        raise NotImplementedError

    def ineq_egrad(self, i, x):
        raise NotImplementedError

    def ineq_ehess(self, i, x, v):
        raise NotImplementedError

    # --- derived exactly as pymanopt.Problem / utils.NonlinearProblem do ---
    def preconditioner(self, x, v):
        return v

    def riemannian_gradient(self, x):
        return self.manifold.euclidean_to_riemannian_gradient(x, self.euclidean_gradient(x))

    def riemannian_hessian(self, x, v):
        return self.manifold.euclidean_to_riemannian_hessian(
            x, self.euclidean_gradient(x), self.euclidean_hessian(x, v), v)

    @property
    def ineqconstraints_all(self):
        return [(lambda x, i=i: self.ineq(i, x)) for i in range(self.num_ineqconstraints)]

    @property
    def ineqconstraints_euclidean_gradient_all(self):
        return [(lambda x, i=i: self.ineq_egrad(i, x)) for i in range(self.num_ineqconstraints)]

    @property
    def ineqconstraints_euclidean_hessian_all(self):
        return [(lambda x, v, i=i: self.ineq_ehess(i, x, v)) for i in range(self.num_ineqconstraints)]

    @property
    def ineqconstraints_riemannian_gradient_all(self):
        # utils.py:104-113
        man = self.manifold
        return [
            (lambda x, i=i: man.euclidean_to_riemannian_gradient(x, self.ineq_egrad(i, x)))
            for i in range(self.num_ineqconstraints)
        ]

    @property
    def ineqconstraints_riemannian_hessian_all(self):
        # utils.py:157-169
        man = self.manifold
        return [
            (lambda x, v, i=i: man.euclidean_to_riemannian_hessian(
                x, self.ineq_egrad(i, x), self.ineq_ehess(i, x, v), v))
            for i in range(self.num_ineqconstraints)
        ]


class NonnegPCAProblem(OracleProblem):
    """min -x'Zx on Sphere(n), x_i >= 0  (src/NonnegPCA/coordinator.py:37-95).

    Z is the square, NON-symmetric matrix of dataset/NonnegPCA/<inst>/Z.csv used
    directly in the quadratic form (:52-54); constraints g_i = -x_i (:66-70).
    """

    family = "nonnegpca_sphere"

    def __init__(self, Z, x0, y0=None, eps=0.0):
        Z = np.asarray(Z, dtype=float)
        n = Z.shape[0]
        self.Z = Z
        self.eps = float(eps)
        super().__init__(M.Sphere(n), np.asarray(x0, dtype=float),
                         np.ones(n) if y0 is None else y0)

    def cost(self, x):
        return -x @ self.Z @ x

    def euclidean_gradient(self, x):
        return -(x @ self.Z) - self.Z @ x

    def euclidean_hessian(self, x, v):
        return -(v @ self.Z) - self.Z @ v

    def ineq(self, i, x):
        return -x[i] - self.eps

    def ineq_egrad(self, i, x):
        e = np.zeros_like(x)
        e[i] = -1.0
        return e

    def ineq_ehess(self, i, x, v):
        return np.zeros_like(x)

    @staticmethod
    def manviofun(problem, x):
        # src/NonnegPCA/simulator.py:12-14
        return np.linalg.norm(x) - 1


class NonnegPCAStiefelProblem(OracleProblem):
    """EXTRAPOLATED (no reference code).

    f = -tr(X'ZX) on Stiefel(n,p) (`manifold='stiefel'`, offset constraints
    g_ij = -X_ij - eps as in src/Rosenbrock/coordinator.py:62) or on
    Oblique(n,p) (`manifold='oblique'`: p unit-norm columns sharing Z, which
    decouples into p reference-exact Sphere problems).  Constraint index
    i <-> (row, col) in row-major order of X.
    """

    family = "nonnegpca_matrix"

    def __init__(self, Z, X0, y0=None, eps=0.01, manifold="stiefel", closed_form=False):
        Z = np.asarray(Z, dtype=float)
        X0 = np.asarray(X0, dtype=float)
        n, p = X0.shape
        self.Z, self.eps, self.n, self.p = Z, float(eps), n, p
        man = M.Stiefel(n, p) if manifold == "stiefel" else M.Oblique(n, p)
        super().__init__(man, X0, np.ones(n * p) if y0 is None else y0)
        # closed_form: the aggregate operators of riptrm_oracle.py in matrix form instead of
        # m = n p per-constraint closures; checked against the per-constraint path in tests/test_oracle_stiefel.py
        self.closed = _StiefelClosedForm(self) if (closed_form and manifold == "stiefel") else None

    def cost(self, X):
        return -np.trace(X.T @ self.Z @ X)

    def euclidean_gradient(self, X):
        return -(self.Z.T @ X) - self.Z @ X

    def euclidean_hessian(self, X, V):
        return -(self.Z.T @ V) - self.Z @ V

    def ineq(self, i, X):
        return -X.reshape(-1)[i] - self.eps

    def ineq_egrad(self, i, X):
        e = np.zeros(self.n * self.p)
        e[i] = -1.0
        return e.reshape(self.n, self.p)

    def ineq_ehess(self, i, X, V):
        return np.zeros_like(X)

    @staticmethod
    def manviofun(problem, X):
        if isinstance(problem.manifold, M.Oblique):
            return np.linalg.norm(np.linalg.norm(X, axis=0) - 1)
        return np.linalg.norm(X.T @ X - np.eye(X.shape[1]))


class _StiefelClosedForm:
    """Matrix forms of the RIPTRM operators for NonnegPCAStiefelProblem: g_ij = -X_ij - eps, so
    egrad g_ij = -E_ij, ehess g_ij = 0, grad s_ij = P_X(E_ij), and every sum over constraints is one projection."""

    def __init__(self, problem):
        self.pb = problem
        self.S = problem.Z + problem.Z.T

    def _Y(self, y):
        return np.asarray(y).reshape(self.pb.n, self.pb.p)

    def slack(self, X):
        return (X + self.pb.eps).reshape(-1)

    def grad_lagrangian(self, X, y):                      # grad f + sum y_i grad g_i
        man = self.pb.manifold
        return man.projection(X, -self.S @ X) + man.projection(X, -self._Y(y))

    def hess_lagrangian(self, X, y, V):                   # Hess f[V] + sum y_i Hess g_i[V]
        man = self.pb.manifold
        hf = man.euclidean_to_riemannian_hessian(X, -self.S @ X, -self.S @ V, V)
        hg = man.euclidean_to_riemannian_hessian(X, -self._Y(y), np.zeros_like(X), V)
        return hf + hg

    def G_apply(self, X, w):                              # sum w_i grad s_i
        return self.pb.manifold.projection(X, self._Y(w))

    def Gadj_apply(self, X, V, euclidean_embedded):       # <grad s_i, V>
        return (V if euclidean_embedded else self.pb.manifold.projection(X, V)).reshape(-1)


class RosenbrockProblem(OracleProblem):
    """Quadratic chain cost on Grassmann(n,k), vec(X)_i >= -0.01
    (src/Rosenbrock/coordinator.py:33-91; config_simulation.yaml:10-12: n=5,k=3,alpha=1e7).

    cost (:44-51): v = X.flatten() (row-major); sum_{i<len-1} alpha (v[i+1]-v[i])^2 + (1-v[i])^2.
    constraints (:58-63): g_i = -v[i] - 0.01.  x0 = I[:, :k] (:78-84), y0 = 1 (:87-91).
    """

    family = "rosenbrock_grassmann"

    def __init__(self, n=5, k=3, alpha=1e7):
        self.n, self.k, self.alpha = n, k, float(alpha)
        x0 = np.abs(np.eye(n)[:, :k])
        super().__init__(M.Grassmann(n, k), x0, np.ones(n * k))

    def cost(self, X):
        v = X.flatten()
        val = 0
        for i in range(len(v) - 1):
            val = val + self.alpha * (v[i + 1] - v[i]) ** 2 + (1 - v[i]) ** 2
        return val

    def _apply_T(self, v):
        # Hessian of the cost (constant, tridiagonal) applied to a flat vector
        a = self.alpha
        out = np.zeros_like(v)
        d = v[1:] - v[:-1]
        out[:-1] += -2 * a * d + 2 * v[:-1]
        out[1:] += 2 * a * d
        return out

    def euclidean_gradient(self, X):
        v = X.flatten()
        g = self._apply_T(v)
        g[:-1] += -2.0
        return g.reshape(X.shape)

    def euclidean_hessian(self, X, V):
        return self._apply_T(V.flatten()).reshape(X.shape)

    def ineq(self, i, X):
        return -X.flatten()[i] - 0.01

    def ineq_egrad(self, i, X):
        e = np.zeros(X.size)
        e[i] = -1.0
        return e.reshape(X.shape)

    def ineq_ehess(self, i, X, V):
        return np.zeros_like(X)

    @staticmethod
    def manviofun(problem, x):
        # src/Rosenbrock/simulator.py:107-114
        return 0 if np.linalg.matrix_rank(x) == problem.manifold._p else np.inf


class StableIdentificationProblem(OracleProblem):
    """Fit A=(J-R)Q on Product[Skew(d),SPD(d),SPD(d)] with box constraints on entries of A
    (src/StableIdentification/coordinator.py:34-179; config_simulation.yaml:10-12).

    cost (:92-98): E = XP - (I + h A) X, f = tr(E E')/N.
    constraints (:108-152) from `constset` rows [type,row,col,a,b,(unused)]:
      type 0/1 -> two one-box constraints  g = -A[r,c] + a  and  g = A[r,c] - b;
      type 2   -> one two-box constraint   g = -(A[r,c]-a)^2 + b^2.
    """

    family = "stableid_product"

    def __init__(self, X, XP, h, constset, x0, y0):
        X = np.asarray(X, dtype=float)
        XP = np.asarray(XP, dtype=float)
        d = X.shape[0]
        self.X, self.XP, self.h, self.d, self.N = X, XP, float(h), d, X.shape[1]
        self.constset = np.asarray(constset, dtype=float)
        # (kind, row, col, a): kind 0: -A+a ; 1: A-a ; 2: -(A-a)^2+b^2 (b in slot 4)
        spec = []
        for row in self.constset:
            t, r, c = int(row[0]), int(row[1]), int(row[2])
            if t in (0, 1):
                spec.append((0, r, c, row[3], 0.0))
                spec.append((1, r, c, row[4], 0.0))
            elif t == 2:
                spec.append((2, r, c, row[3], row[4] ** 2))
            else:
                raise ValueError("Invalid constraint type")
        self.spec = spec
        man = M.Product([M.SkewSymmetric(d), M.SymmetricPositiveDefinite(d), M.SymmetricPositiveDefinite(d)])
        super().__init__(man, [np.asarray(a, dtype=float) for a in x0], y0)
        assert len(spec) == self.num_ineqconstraints

    @classmethod
    def from_dataset(cls, path, initialpoint="a", Xset=(1, 2, 3, 4, 5), h=0.02, is_X_noisy=True):
        Xs, XPs = [], []
        for k in Xset:
            Xo = np.loadtxt(f"{path}/{'noisyX' if is_X_noisy else 'X'}_{k}.csv")
            Xs.append(Xo[:, :-1])
            XPs.append(Xo[:, 1:])
        x0 = [np.loadtxt(f"{path}/init{c}_{initialpoint}.csv") for c in "JRQ"]
        return cls(np.hstack(Xs), np.hstack(XPs), h, np.loadtxt(f"{path}/constset.csv"),
                   x0, np.loadtxt(f"{path}/initineqLagmult.csv"))

    def _A(self, x):
        J, R, Q = x
        return (J - R) @ Q

    def cost(self, x):
        A = self._A(x)
        E = self.XP - (np.eye(self.d) + self.h * A) @ self.X
        return np.trace(E @ E.T) / self.N

    @staticmethod
    def _pull(x, PhiA):
        J, R, Q = x
        gJ = PhiA @ Q.T
        return [gJ, -gJ, (J - R).T @ PhiA]

    @staticmethod
    def _pull_d(x, v, PhiA, dPhiA):
        J, R, Q = x
        dJ, dR, dQ = v
        hJ = dPhiA @ Q.T + PhiA @ dQ.T
        return [hJ, -hJ, (dJ - dR).T @ PhiA + (J - R).T @ dPhiA]

    def _GA(self, x):
        A = self._A(x)
        E = self.XP - (np.eye(self.d) + self.h * A) @ self.X
        return -2 * self.h * (E @ self.X.T) / self.N

    def euclidean_gradient(self, x):
        return self._pull(x, self._GA(x))

    def euclidean_hessian(self, x, v):
        J, R, Q = x
        dJ, dR, dQ = v
        dA = (dJ - dR) @ Q + (J - R) @ dQ
        dGA = 2 * self.h ** 2 * (dA @ (self.X @ self.X.T)) / self.N
        return self._pull_d(x, v, self._GA(x), dGA)

    def ineq(self, i, x):
        kind, r, c, a, b2 = self.spec[i]
        Arc = self._A(x)[r, c]
        if kind == 0:
            return -Arc + a
        if kind == 1:
            return Arc - a
        return -(Arc - a) ** 2 + b2

    def _Phi(self, i, x):
        kind, r, c, a, b2 = self.spec[i]
        P = np.zeros((self.d, self.d))
        if kind == 0:
            P[r, c] = -1.0
        elif kind == 1:
            P[r, c] = 1.0
        else:
            P[r, c] = -2 * (self._A(x)[r, c] - a)
        return P

    def ineq_egrad(self, i, x):
        return self._pull(x, self._Phi(i, x))

    def ineq_ehess(self, i, x, v):
        kind, r, c, a, b2 = self.spec[i]
        dP = np.zeros((self.d, self.d))
        if kind == 2:
            J, R, Q = x
            dJ, dR, dQ = v
            dA = (dJ - dR) @ Q + (J - R) @ dQ
            dP[r, c] = -2 * dA[r, c]
        return self._pull_d(x, v, self._Phi(i, x), dP)

    @staticmethod
    def manviofun(problem, x):
        # src/StableIdentification/simulator.py:11-33
        J, R, Q = x
        manvio = np.linalg.norm(J + J.T) + np.linalg.norm(R - R.T) + np.linalg.norm(Q - Q.T)
        if not np.all(np.linalg.eigvalsh(R) > 0) or not np.all(np.linalg.eigvalsh(Q) > 0):
            manvio = np.inf
        return manvio


# ---------------------------------------------------------------------------
# Synthetic instance generators following src/NonnegPCA/generator.py:9-65
# (the reference is unseeded; here seed = instance id).
# ---------------------------------------------------------------------------
def nonnegpca_generate_Z(dim, snr=0.5, delta=0.7, seed=0):
    rs = np.random.RandomState(seed)
    samplesize = int(np.floor(delta * dim))
    S = rs.choice(dim, samplesize, replace=False)
    v = np.zeros(dim)
    v[S] = 1 / np.sqrt(samplesize)
    Z = np.sqrt(snr) * np.outer(v, v)
    noise = rs.randn(dim, dim) / np.sqrt(dim)
    for ii in range(dim):
        noise[ii, ii] = rs.randn() * 2 / np.sqrt(dim)
    return Z + noise, rs


def nonnegpca_generate_instance(dim=50, snr=0.5, delta=0.7, seed=0):
    """(Z, x0, y0) by the reference generator law: Z (generator.py:9-31), feasible x0 (:46-51), y0=1 (:63)."""
    Z, rs = nonnegpca_generate_Z(dim, snr, delta, seed)
    x0 = rs.rand(dim)
    x0 = np.abs(x0 / np.linalg.norm(x0))
    return Z, x0, np.ones(dim)


def nonnegpca_generate_sweep(first_instance, instances, points_per_instance, dim=50):
    """`instances` x `points_per_instance` (instance, initialpoint) pairs, instance-major: point 0 of an instance is
    the generator's x0 (seed = instance id), points k >= 1 follow generator.py:46-51 from seed 2e9 + 16*id + k (k < 16).
    Returns Z [instances, dim, dim], x0 / y0 [pairs, dim]."""
    pairs = instances * points_per_instance
    Z, x0, y0 = np.empty((instances, dim, dim)), np.empty((pairs, dim)), np.ones((pairs, dim))
    for i in range(instances):
        inst = first_instance + i
        Z[i], x0[i * points_per_instance], _ = nonnegpca_generate_instance(dim, seed=inst)
        for k in range(1, points_per_instance):
            u = np.random.RandomState(2000000000 + 16 * inst + k).rand(1, dim)
            x0[i * points_per_instance + k] = np.abs(u / np.linalg.norm(u, axis=1, keepdims=True))[0]
    return Z, x0, y0
