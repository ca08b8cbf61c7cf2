"""Deterministic orthonormal tangent bases -- the `basisfun` option of the reference (src/solver/RIPTRM.py:341,
called at :429 and :600 as `basisfun(manifold, x)` -> list of manifold.dim tangent vectors).

The reference's default is `tangentorthobasis` (src/solver/utils.py:388-397): Gram-Schmidt on RANDOM tangent vectors, so
two runs of the reference itself use different bases.  The trust-region step and the smallest eigenvalue of the
representation matrix do not depend on which orthonormal basis is used (only their rounding does), so the CUDA path
builds a fixed one per manifold (csrc/fam_*.cuh `coord_setup / from_coords / to_coords`); the functions below are the
same constructions in NumPy, to be passed as `option['basisfun']` when a reference run should be reproducible (the
golden runs of tests/golden/make_golden.py do).

  Sphere(n)            columns 2..n of the Householder reflector that maps e_1 to -sign(x_1) x
  Grassmann(n, p)      X_perp E_ab, X_perp = the last n-p columns of the complete QR factor of X
  SkewSymmetric(d)     (E_ab - E_ba) / sqrt(2), a < b
  SPD(d) at P = L L'   L E L' with E = E_aa or (E_ab + E_ba) / sqrt(2): orthonormal in tr(P^-1 A P^-1 B)
  Product              the component bases, each padded with zero components
"""
import numpy as np


def sphere_basis(x):
    x = np.asarray(x, dtype=float).reshape(-1)
    n = x.size
    sigma = 1.0 if x[0] >= 0 else -1.0
    u = x.copy()
    u[0] += sigma
    beta = 1.0 / (1.0 + abs(x[0]))
    basis = []
    for j in range(1, n):
        b = -beta * u[j] * u
        b[j] += 1.0
        basis.append(b)
    return basis


def grassmann_basis(X):
    X = np.asarray(X, dtype=float)
    n, p = X.shape
    Q, _ = np.linalg.qr(X, mode="complete")
    perp = Q[:, p:]
    basis = []
    for a in range(n - p):
        for b in range(p):
            K = np.zeros((n - p, p))
            K[a, b] = 1.0
            basis.append(perp @ K)
    return basis


def skew_basis(d):
    basis = []
    for a in range(d):
        for b in range(a + 1, d):
            E = np.zeros((d, d))
            E[a, b] = 1.0 / np.sqrt(2.0)
            E[b, a] = -1.0 / np.sqrt(2.0)
            basis.append(E)
    return basis


def spd_basis(P):
    P = np.asarray(P, dtype=float)
    d = P.shape[0]
    L = np.linalg.cholesky(P)
    basis = []
    for a in range(d):
        E = np.zeros((d, d))
        E[a, a] = 1.0
        basis.append(L @ E @ L.T)
    for a in range(d):
        for b in range(a + 1, d):
            E = np.zeros((d, d))
            E[a, b] = E[b, a] = 1.0 / np.sqrt(2.0)
            basis.append(L @ E @ L.T)
    return basis


def _component_basis(manifold, x):
    name = type(manifold).__name__
    if name == "Sphere":
        shape = np.asarray(x).shape
        return [b.reshape(shape) for b in sphere_basis(x)]
    if name == "Grassmann":
        return grassmann_basis(x)
    if name == "SkewSymmetric":
        return skew_basis(np.asarray(x).shape[0])
    if name == "SymmetricPositiveDefinite":
        return spd_basis(x)
    raise NotImplementedError(f"no deterministic tangent basis for manifold {name}")


def deterministic_basisfun(manifold, x):
    """`basisfun(manifold, x)` for the manifolds of the reference's three workloads."""
    if type(manifold).__name__ == "Product":
        comps = getattr(manifold, "manifolds", None) or getattr(manifold, "_manifolds")
        zero = manifold.zero_vector(x)
        cls = type(zero)
        basis = []
        for k, (mk, xk) in enumerate(zip(comps, x)):
            for b in _component_basis(mk, xk):
                v = [np.zeros_like(np.asarray(z, dtype=float)) for z in zero]
                v[k] = b
                basis.append(cls(v) if cls is not list else v)
        return basis
    return _component_basis(manifold, x)
