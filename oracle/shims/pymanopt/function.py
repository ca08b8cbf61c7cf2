"""Shim of pymanopt.function: the `autograd` backend decorator, on oracle/shims/adlite.py."""
from oracle.shims import adlite


class Function:
    def __init__(self, function, manifold):
        self._function = function
        self._manifold = manifold
        layout = manifold.point_layout
        self._nargs = len(layout) if isinstance(layout, (tuple, list)) else 1

    def __call__(self, *args):
        return self._function(*args)

    def get_gradient_operator(self):
        fn, nargs = self._function, self._nargs

        def gradient_operator(*args):
            g = adlite.gradient(fn, args)
            return g[0] if nargs == 1 else g
        return gradient_operator

    def get_hessian_operator(self):
        fn, nargs = self._function, self._nargs

        def hessian_operator(*args):
            points, vectors = args[:nargs], args[nargs:]
            h = adlite.hessian_vector_product(fn, points, vectors)
            return h[0] if nargs == 1 else h
        return hessian_operator


def autograd(manifold):
    def decorator(function):
        return Function(function, manifold)
    return decorator
