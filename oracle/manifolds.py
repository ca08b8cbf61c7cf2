"""TEST INFRASTRUCTURE -- CPU oracle, not product code.

NumPy restatement of the pymanopt 2.x manifold classes the reference calls on
the RIPTRM hot path.  pymanopt is an un-vendored, un-pinned third-party
dependency of the reference (Dockerfile:12 ``pip install pymanopt``; the API
names the reference uses -- ``euclidean_to_riemannian_gradient``,
``to_tangent_space``, ``pymanopt.function.autograd(manifold)`` -- imply
pymanopt >= 2.0, most likely 2.2.x).  It is absent from /root/reference and
from this image, so its published formulas are restated here.

Reference call sites that fix which methods are needed
(src/solver/RIPTRM.py): inner_product :44, zero_vector :47, to_tangent_space
:210, dim :447, euclidean_to_riemannian_gradient :482/:545,
euclidean_to_riemannian_hessian :517, norm :593/:735, retraction :744,
typical_dist :857; (src/solver/utils.py): dist :349, norm :292.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
import numpy as np


def multitransp(A):
    return np.swapaxes(A, -1, -2) if A.ndim >= 2 else A


def multisym(A):
    return 0.5 * (A + multitransp(A))


def multiskew(A):
    return 0.5 * (A - multitransp(A))


class _TangentList(list):
    """pymanopt.manifolds.product._ProductTangentVector: list with vector algebra."""

    # pymanopt.tools.ndarraySequenceMixin: keep numpy scalars from vectorising over the list
    __array_priority__ = 1000
    __array_ufunc__ = None

    def __add__(self, other):
        assert len(self) == len(other)
        return _TangentList([v + other[k] for k, v in enumerate(self)])

    def __sub__(self, other):
        assert len(self) == len(other)
        return _TangentList([v - other[k] for k, v in enumerate(self)])

    def __mul__(self, other):
        return _TangentList([other * val for val in self])

    __rmul__ = __mul__

    def __truediv__(self, other):
        return _TangentList([val / other for val in self])

    def __neg__(self):
        return _TangentList([-val for val in self])


class Manifold:
    point_layout = 1
    name = "manifold"

    def to_tangent_space(self, point, vector):
        return self.projection(point, vector)

    def euclidean_to_riemannian_gradient(self, point, euclidean_gradient):
        return self.projection(point, euclidean_gradient)

    def embedding(self, point, tangent_vector):
        return tangent_vector


class Sphere(Manifold):
    """pymanopt.manifolds.sphere.Sphere (unit Frobenius-norm arrays of a shape)."""

    def __init__(self, *shape):
        self._shape = tuple(shape)
        self.dim = int(np.prod(shape)) - 1
        self.typical_dist = np.pi
        self.name = f"Sphere{self._shape}"

    def inner_product(self, point, a, b):
        return np.tensordot(a, b, axes=a.ndim)

    def norm(self, point, v):
        return np.linalg.norm(v)

    def dist(self, a, b):
        inner = max(min(self.inner_product(a, a, b), 1), -1)
        return np.arccos(inner)

    def projection(self, point, vector):
        return vector - self.inner_product(point, point, vector) * point

    to_tangent_space = projection
    euclidean_to_riemannian_gradient = projection

    def euclidean_to_riemannian_hessian(self, point, egrad, ehess, tv):
        return self.projection(point, ehess) - self.inner_product(point, point, egrad) * tv

    def retraction(self, point, tv):
        a = point + tv
        return a / np.linalg.norm(a)

    def zero_vector(self, point):
        return np.zeros(self._shape)

    def random_point(self, rng=np.random):
        a = rng.normal(size=self._shape)
        return a / np.linalg.norm(a)

    def random_tangent_vector(self, point, rng=np.random):
        v = self.projection(point, rng.normal(size=self._shape))
        return v / np.linalg.norm(v)


class Grassmann(Manifold):
    """pymanopt.manifolds.grassmann.Grassmann(n, p) (k=1), orthonormal n x p representatives."""

    def __init__(self, n, p):
        self._n, self._p = n, p
        self.dim = n * p - p * p
        self.typical_dist = np.sqrt(p)
        self.name = f"Grassmann({n},{p})"

    def inner_product(self, point, a, b):
        return np.tensordot(a, b, axes=a.ndim)

    def norm(self, point, v):
        return np.linalg.norm(v)

    def dist(self, a, b):
        s = np.linalg.svd(multitransp(a) @ b, compute_uv=False)
        s[s > 1] = 1
        s = np.arccos(s)
        return np.linalg.norm(s)

    def projection(self, point, vector):
        return vector - point @ (multitransp(point) @ vector)

    to_tangent_space = projection
    euclidean_to_riemannian_gradient = projection

    def euclidean_to_riemannian_hessian(self, point, egrad, ehess, tv):
        PXehess = self.projection(point, ehess)
        XtG = multitransp(point) @ egrad
        HXtG = tv @ XtG
        return PXehess - HXtG

    def retraction(self, point, tv):
        u, _, vt = np.linalg.svd(point + tv, full_matrices=False)
        return u @ vt

    def zero_vector(self, point):
        return np.zeros((self._n, self._p))

    def random_point(self, rng=np.random):
        q, _ = np.linalg.qr(rng.normal(size=(self._n, self._p)))
        return q

    def random_tangent_vector(self, point, rng=np.random):
        v = self.projection(point, rng.normal(size=point.shape))
        return v / np.linalg.norm(v)


class Stiefel(Manifold):
    """pymanopt.manifolds.stiefel.Stiefel(n, p) (k=1); default retraction 'qr'."""

    def __init__(self, n, p, retraction="qr"):
        self._n, self._p = n, p
        self.dim = int(n * p - p * (p + 1) / 2)
        self.typical_dist = np.sqrt(p)
        self._retr = retraction
        self.name = f"Stiefel({n},{p})"

    def inner_product(self, point, a, b):
        return np.tensordot(a, b, axes=a.ndim)

    def norm(self, point, v):
        return np.linalg.norm(v)

    def dist(self, a, b):
        raise NotImplementedError("pymanopt's Stiefel has no dist")

    def projection(self, point, vector):
        return vector - point @ multisym(multitransp(point) @ vector)

    to_tangent_space = projection
    euclidean_to_riemannian_gradient = projection

    def euclidean_to_riemannian_hessian(self, point, egrad, ehess, tv):
        XtG = multitransp(point) @ egrad
        symXtG = multisym(XtG)
        HsymXtG = tv @ symXtG
        return self.projection(point, ehess - HsymXtG)

    def retraction(self, point, tv):
        a = point + tv
        if self._retr == "qr":
            q, r = np.linalg.qr(a)
            return q * np.sign(np.sign(np.diag(r)) + 0.5)
        u, _, vt = np.linalg.svd(a, full_matrices=False)
        return u @ vt

    def zero_vector(self, point):
        return np.zeros((self._n, self._p))

    def random_point(self, rng=np.random):
        q, _ = np.linalg.qr(rng.normal(size=(self._n, self._p)))
        return q

    def random_tangent_vector(self, point, rng=np.random):
        v = self.projection(point, rng.normal(size=point.shape))
        return v / np.linalg.norm(v)


class Oblique(Manifold):
    """pymanopt.manifolds.oblique.Oblique(m, n): m x n matrices with unit-norm columns."""

    def __init__(self, m, n):
        self._m, self._n = m, n
        self.dim = (m - 1) * n
        self.typical_dist = np.pi * np.sqrt(n)
        self.name = f"Oblique({m},{n})"

    def inner_product(self, point, a, b):
        return np.tensordot(a, b, axes=a.ndim)

    def norm(self, point, v):
        return np.linalg.norm(v)

    def dist(self, a, b):
        return np.linalg.norm(np.arccos(np.clip((a * b).sum(0), -1, 1)))

    def projection(self, point, vector):
        return vector - point * ((vector * point).sum(0))

    to_tangent_space = projection
    euclidean_to_riemannian_gradient = projection

    def euclidean_to_riemannian_hessian(self, point, egrad, ehess, tv):
        PXehess = self.projection(point, ehess)
        return PXehess - tv * ((point * egrad).sum(0))

    def retraction(self, point, tv):
        a = point + tv
        return a / np.linalg.norm(a, axis=0)[np.newaxis, :]

    def zero_vector(self, point):
        return np.zeros((self._m, self._n))

    def random_point(self, rng=np.random):
        a = rng.normal(size=(self._m, self._n))
        return a / np.linalg.norm(a, axis=0)[np.newaxis, :]

    def random_tangent_vector(self, point, rng=np.random):
        v = self.projection(point, rng.normal(size=point.shape))
        return v / np.linalg.norm(v)


class _Euclidean(Manifold):
    def __init__(self, shape, dim):
        self._shape = shape
        self.dim = dim
        self.typical_dist = np.sqrt(dim)

    def inner_product(self, point, a, b):
        return float(np.real(np.tensordot(a.conj(), b, axes=a.ndim)))

    def norm(self, point, v):
        return np.linalg.norm(v)

    def dist(self, a, b):
        return np.linalg.norm(a - b)

    def projection(self, point, vector):
        return vector

    def euclidean_to_riemannian_hessian(self, point, egrad, ehess, tv):
        return ehess

    def retraction(self, point, tv):
        return point + tv

    def zero_vector(self, point):
        return np.zeros(self._shape)


class SkewSymmetric(_Euclidean):
    """pymanopt.manifolds.euclidean.SkewSymmetric(n) (k=1)."""

    def __init__(self, n):
        super().__init__((n, n), int(n * (n - 1) / 2))
        self._n = n
        self.name = f"SkewSymmetric({n})"

    def projection(self, point, vector):
        return multiskew(vector)

    def euclidean_to_riemannian_hessian(self, point, egrad, ehess, tv):
        return multiskew(ehess)

    def random_point(self, rng=np.random):
        return multiskew(rng.normal(size=self._shape))

    def random_tangent_vector(self, point, rng=np.random):
        v = self.random_point(rng)
        return multiskew(v / self.norm(point, v))


class SymmetricPositiveDefinite(Manifold):
    """pymanopt.manifolds.positive_definite.SymmetricPositiveDefinite(n) (k=1), affine-invariant metric.

    VERSION-SENSITIVE: the retraction is the second-order
    one of pymanopt >= 2.1 (``sym(P + V + V P^-1 V / 2)``); pymanopt 2.0 used exp.
    """

    def __init__(self, n):
        self._n = n
        self.dim = int(n * (n + 1) / 2)
        self.typical_dist = np.sqrt(self.dim)
        self.name = f"SymmetricPositiveDefinite({n})"

    def inner_product(self, point, a, b):
        p_inv_a = np.linalg.solve(point, a)
        p_inv_b = p_inv_a if a is b else np.linalg.solve(point, b)
        return np.tensordot(p_inv_a, multitransp(p_inv_b), axes=a.ndim)

    def norm(self, point, v):
        return np.sqrt(self.inner_product(point, v, v))

    def dist(self, a, b):
        c = np.linalg.cholesky(a)
        c_inv = np.linalg.inv(c)
        w = np.linalg.eigvalsh(multisym(c_inv @ b @ multitransp(c_inv)))
        return np.linalg.norm(np.log(w))

    def projection(self, point, vector):
        return multisym(vector)

    to_tangent_space = projection

    def euclidean_to_riemannian_gradient(self, point, egrad):
        return point @ multisym(egrad) @ point

    def euclidean_to_riemannian_hessian(self, point, egrad, ehess, tv):
        return point @ multisym(ehess) @ point + multisym(tv @ multisym(egrad) @ point)

    def retraction(self, point, tv):
        p_inv_tv = np.linalg.solve(point, tv)
        return multisym(point + tv + tv @ p_inv_tv / 2)

    def zero_vector(self, point):
        return np.zeros((self._n, self._n))

    def random_point(self, rng=np.random):
        d = 1.0 + rng.uniform(size=self._n)
        q, _ = np.linalg.qr(rng.normal(size=(self._n, self._n)))
        return multisym(q @ np.diag(d) @ q.T)

    def random_tangent_vector(self, point, rng=np.random):
        v = multisym(rng.normal(size=(self._n, self._n)))
        return v / self.norm(point, v)


class Product(Manifold):
    """pymanopt.manifolds.product.Product: points are lists, tangent vectors _TangentList."""

    def __init__(self, manifolds):
        self.manifolds = tuple(manifolds)
        self.dim = int(np.sum([m.dim for m in manifolds]))
        self.point_layout = tuple(m.point_layout for m in manifolds)
        self.name = "Product[" + ",".join(m.name for m in manifolds) + "]"

    @property
    def typical_dist(self):
        return np.sqrt(np.sum([m.typical_dist ** 2 for m in self.manifolds]))

    def inner_product(self, point, a, b):
        return np.sum([m.inner_product(point[k], a[k], b[k]) for k, m in enumerate(self.manifolds)])

    def norm(self, point, v):
        return np.sqrt(self.inner_product(point, v, v))

    def dist(self, a, b):
        return np.sqrt(np.sum([m.dist(a[k], b[k]) ** 2 for k, m in enumerate(self.manifolds)]))

    def projection(self, point, vector):
        return _TangentList([m.projection(point[k], vector[k]) for k, m in enumerate(self.manifolds)])

    def to_tangent_space(self, point, vector):
        return _TangentList([m.to_tangent_space(point[k], vector[k]) for k, m in enumerate(self.manifolds)])

    def euclidean_to_riemannian_gradient(self, point, egrad):
        return _TangentList(
            [m.euclidean_to_riemannian_gradient(point[k], egrad[k]) for k, m in enumerate(self.manifolds)]
        )

    def euclidean_to_riemannian_hessian(self, point, egrad, ehess, tv):
        return _TangentList(
            [
                m.euclidean_to_riemannian_hessian(point[k], egrad[k], ehess[k], tv[k])
                for k, m in enumerate(self.manifolds)
            ]
        )

    def retraction(self, point, tv):
        return [m.retraction(point[k], tv[k]) for k, m in enumerate(self.manifolds)]

    def zero_vector(self, point):
        return _TangentList([m.zero_vector(point[k]) for k, m in enumerate(self.manifolds)])

    def embedding(self, point, tv):
        return _TangentList([m.embedding(point[k], tv[k]) for k, m in enumerate(self.manifolds)])

    def random_point(self, rng=np.random):
        return [m.random_point(rng) for m in self.manifolds]

    def random_tangent_vector(self, point, rng=np.random):
        scale = len(self.manifolds) ** (-1 / 2)
        return _TangentList(
            [scale * m.random_tangent_vector(point[k], rng) for k, m in enumerate(self.manifolds)]
        )
