"""TEST INFRASTRUCTURE -- stand-in for the `pymanopt` package (absent from this image).

Restates the small part of pymanopt 2.x's public surface the reference imports
(src/solver/utils.py:33-52,95-97,147-150; src/solver/RIPTRM.py:12,449;
coordinators' `pymanopt.manifolds.*` and `@pymanopt.function.autograd`), so the
UNMODIFIED reference can run in this container and produce golden vectors
(tests/golden/make_golden.py).  Manifold formulas live in oracle/manifolds.py.
"""
import functools

from . import manifolds, function, tools  # noqa: F401

__version__ = "2.2-shim"


class Problem:
    """pymanopt.core.problem.Problem (cost + lazily derived gradient/Hessian operators)."""

    def __init__(self, manifold, cost, *, preconditioner=None):
        self.manifold = manifold
        self._original_cost = cost
        self._cost = self._wrap_function(cost)
        self._euclidean_gradient = None
        self._riemannian_gradient = None
        self._euclidean_hessian = None
        self._riemannian_hessian = None
        if preconditioner is None:
            def preconditioner(point, tangent_vector):
                return tangent_vector
        self.preconditioner = preconditioner

    @staticmethod
    def _flatten_arguments(arguments, signature):
        assert len(arguments) == len(signature)
        flat = []
        for i, group_size in enumerate(signature):
            argument = arguments[i]
            if isinstance(group_size, (list, tuple)) or group_size > 1:
                flat.extend(argument)
            elif isinstance(argument, (list, tuple)) and group_size == 1 and False:
                flat.extend(argument)
            else:
                flat.append(argument)
        return flat

    def _unpack_point(self, args):
        """Product points/tangents arrive as lists; flatten them into positional arguments."""
        layout = self.manifold.point_layout
        if isinstance(layout, (tuple, list)):
            flat = []
            for a in args:
                flat.extend(list(a))
            return flat
        return list(args)

    def _group(self, values):
        layout = self.manifold.point_layout
        if isinstance(layout, (tuple, list)):
            return list(values)
        return values

    def _wrap_function(self, fn):
        layout = self.manifold.point_layout
        if isinstance(layout, (tuple, list)):
            @functools.wraps(fn)
            def unpack_arguments(*args):
                return fn(*self._unpack_point(args))
            return unpack_arguments
        return fn

    def _wrap_gradient_operator(self, gradient_operator):
        wrapped = self._wrap_function(gradient_operator)
        layout = self.manifold.point_layout
        if isinstance(layout, (tuple, list)):
            @functools.wraps(wrapped)
            def group_return_values(*args):
                return self._group(wrapped(*args))
            return group_return_values
        return wrapped

    def _wrap_hessian_operator(self, hessian_operator, *, embed_tangent_vectors=False):
        wrapped = self._wrap_function(hessian_operator)
        layout = self.manifold.point_layout
        if isinstance(layout, (tuple, list)):
            inner = wrapped

            @functools.wraps(inner)
            def group_return_values(*args):
                return self._group(inner(*args))
            wrapped = group_return_values
        if embed_tangent_vectors:
            op = wrapped

            @functools.wraps(op)
            def embed(point, tangent_vector):
                return op(point, self.manifold.embedding(point, tangent_vector))
            return embed
        return wrapped

    @property
    def cost(self):
        return self._cost

    @property
    def euclidean_gradient(self):
        if self._euclidean_gradient is None:
            self._euclidean_gradient = self._wrap_gradient_operator(
                self._original_cost.get_gradient_operator())
        return self._euclidean_gradient

    @property
    def riemannian_gradient(self):
        if self._riemannian_gradient is None:
            def riemannian_gradient(point):
                return self.manifold.euclidean_to_riemannian_gradient(
                    point, self.euclidean_gradient(point))
            self._riemannian_gradient = riemannian_gradient
        return self._riemannian_gradient

    @property
    def euclidean_hessian(self):
        if self._euclidean_hessian is None:
            self._euclidean_hessian = self._wrap_hessian_operator(
                self._original_cost.get_hessian_operator(), embed_tangent_vectors=True)
        return self._euclidean_hessian

    @property
    def riemannian_hessian(self):
        if self._riemannian_hessian is None:
            def riemannian_hessian(point, tangent_vector):
                return self.manifold.euclidean_to_riemannian_hessian(
                    point,
                    self.euclidean_gradient(point),
                    self.euclidean_hessian(point, tangent_vector),
                    tangent_vector,
                )
            self._riemannian_hessian = riemannian_hessian
        return self._riemannian_hessian
