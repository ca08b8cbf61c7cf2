"""TEST INFRASTRUCTURE -- stand-in for the `autograd` package (absent from this image).

A minimal exact automatic differentiator, just large enough to differentiate
the cost/constraint closures of the reference's three coordinators
(src/NonnegPCA/coordinator.py:52-54,66-70; src/Rosenbrock/coordinator.py:44-51,
58-63; src/StableIdentification/coordinator.py:92-98,108-130) WITHOUT modifying
them.  Gradients are reverse mode (a tape of `Box` nodes); Hessian-vector
products are forward-over-reverse (the tape runs over `Dual` values), which is
the same mathematical object autograd's reverse-over-reverse
`hessian_vector_product` returns.

Used only by oracle/shims/pymanopt/function.py to generate golden vectors from
the unmodified reference (tests/golden/make_golden.py).
"""
import numpy as np


def _is_dual(a):
    return isinstance(a, Dual)


def _val(a):
    return a.val if isinstance(a, Dual) else a


def _tan(a):
    return a.tan if isinstance(a, Dual) else np.zeros_like(np.asarray(a, dtype=float))


class Dual:
    """value + eps * tangent, first order."""

    __array_priority__ = 2000
    __array_ufunc__ = None

    def __init__(self, val, tan):
        self.val = np.asarray(val, dtype=float)
        self.tan = np.asarray(tan, dtype=float)

    shape = property(lambda self: self.val.shape)
    ndim = property(lambda self: self.val.ndim)

    def __len__(self):
        return len(self.val)

    def __neg__(self):
        return Dual(-self.val, -self.tan)

    def __add__(self, o):
        return Dual(self.val + _val(o), self.tan + _tan(o)) if _is_dual(o) else Dual(self.val + o, self.tan + 0 * np.asarray(o))

    __radd__ = __add__

    def __sub__(self, o):
        return Dual(self.val - o.val, self.tan - o.tan) if _is_dual(o) else Dual(self.val - o, self.tan - 0 * np.asarray(o))

    def __rsub__(self, o):
        return Dual(o - self.val, 0 * np.asarray(o) - self.tan)

    def __mul__(self, o):
        if _is_dual(o):
            return Dual(self.val * o.val, self.val * o.tan + self.tan * o.val)
        return Dual(self.val * o, self.tan * o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        assert not _is_dual(o)
        return Dual(self.val / o, self.tan / o)

    def __pow__(self, k):
        assert isinstance(k, (int, float))
        return Dual(self.val ** k, k * self.val ** (k - 1) * self.tan)

    def __matmul__(self, o):
        if _is_dual(o):
            return Dual(self.val @ o.val, self.val @ o.tan + self.tan @ o.val)
        return Dual(self.val @ o, self.tan @ o)

    def __rmatmul__(self, o):
        return Dual(o @ self.val, o @ self.tan)

    def __getitem__(self, idx):
        return Dual(self.val[idx], self.tan[idx])

    @property
    def T(self):
        return Dual(self.val.T, self.tan.T)

    def reshape(self, *shape):
        return Dual(self.val.reshape(*shape), self.tan.reshape(*shape))

    def flatten(self):
        return Dual(self.val.flatten(), self.tan.flatten())

    def trace(self):
        return Dual(np.trace(self.val), np.trace(self.tan))

    def sum(self):
        return Dual(self.val.sum(), self.tan.sum())


def _outer(a, b):
    """outer product of two 1-D operands that may be Dual."""
    if _is_dual(a) or _is_dual(b):
        av, at, bv, bt = _val(a), _tan(a), _val(b), _tan(b)
        return Dual(np.outer(av, bv), np.outer(av, bt) + np.outer(at, bv))
    return np.outer(a, b)


def _scatter(shape, idx, g):
    if _is_dual(g):
        v = np.zeros(shape)
        t = np.zeros(shape)
        v[idx] = g.val
        t[idx] = g.tan
        return Dual(v, t)
    out = np.zeros(shape)
    out[idx] = g
    return out


def _shape(a):
    return a.shape if hasattr(a, "shape") else ()


def _ndim(a):
    return len(_shape(a))


def _T(a):
    return a.T if _ndim(a) >= 2 else a


def _unbroadcast(g, shape):
    if _shape(g) == tuple(shape):
        return g
    if tuple(shape) == ():
        return g.sum()
    raise NotImplementedError("broadcasting between boxed operands")


class _Tape:
    def __init__(self):
        self.nodes = []


class Box:
    """Reverse-mode tracer.  `value` is an ndarray/float or a Dual."""

    __array_priority__ = 3000
    __array_ufunc__ = None

    def __init__(self, value, tape, parents=()):
        self.value = value
        self.tape = tape
        self.parents = parents  # list of (Box, vjp)
        tape.nodes.append(self)

    shape = property(lambda self: _shape(self.value))
    ndim = property(lambda self: _ndim(self.value))

    def __len__(self):
        return len(self.value)

    @staticmethod
    def _lift(o):
        return o.value if isinstance(o, Box) else o

    def _new(self, value, parents):
        return Box(value, self.tape, parents)

    def __neg__(self):
        return self._new(-self.value, [(self, lambda g: -g)])

    def __add__(self, o):
        if isinstance(o, Box):
            return self._new(
                self.value + o.value,
                [(self, lambda g: _unbroadcast(g, self.shape)), (o, lambda g: _unbroadcast(g, o.shape))],
            )
        return self._new(self.value + o, [(self, lambda g: _unbroadcast(g, self.shape))])

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, Box):
            return self._new(
                self.value - o.value,
                [(self, lambda g: _unbroadcast(g, self.shape)), (o, lambda g: -_unbroadcast(g, o.shape))],
            )
        return self._new(self.value - o, [(self, lambda g: _unbroadcast(g, self.shape))])

    def __rsub__(self, o):
        return self._new(o - self.value, [(self, lambda g: -_unbroadcast(g, self.shape))])

    def __mul__(self, o):
        if isinstance(o, Box):
            a, b = self.value, o.value
            return self._new(
                a * b,
                [(self, lambda g: _unbroadcast(g * b, self.shape)), (o, lambda g: _unbroadcast(g * a, o.shape))],
            )
        return self._new(self.value * o, [(self, lambda g: _unbroadcast(g * o, self.shape))])

    __rmul__ = __mul__

    def __truediv__(self, o):
        assert not isinstance(o, Box)
        return self._new(self.value / o, [(self, lambda g: g / o)])

    def __pow__(self, k):
        a = self.value
        return self._new(a ** k, [(self, lambda g: g * (k * a ** (k - 1)))])

    def __matmul__(self, o):
        a = self.value
        b = self._lift(o)
        out = a @ b
        na, nb = _ndim(a), _ndim(b)
        parents = []
        if na == 2 and nb == 2:
            parents.append((self, lambda g: g @ _T(b)))
            if isinstance(o, Box):
                parents.append((o, lambda g: _T(a) @ g))
        elif na == 1 and nb == 2:
            parents.append((self, lambda g: b @ g))
            if isinstance(o, Box):
                parents.append((o, lambda g: _outer(a, g)))
        elif na == 2 and nb == 1:
            parents.append((self, lambda g: _outer(g, b)))
            if isinstance(o, Box):
                parents.append((o, lambda g: _T(a) @ g))
        elif na == 1 and nb == 1:
            parents.append((self, lambda g: g * b))
            if isinstance(o, Box):
                parents.append((o, lambda g: g * a))
        else:
            raise NotImplementedError
        return self._new(out, parents)

    def __rmatmul__(self, o):
        # o is a plain ndarray here (a Box on the left would have used __matmul__)
        b = self.value
        out = o @ b
        na, nb = _ndim(o), _ndim(b)
        if na == 2 and nb == 2:
            vjp = lambda g: _T(o) @ g
        elif na == 2 and nb == 1:
            vjp = lambda g: _T(o) @ g
        elif na == 1 and nb == 2:
            vjp = lambda g: _outer(o, g)
        else:
            vjp = lambda g: g * o
        return self._new(out, [(self, vjp)])

    def __getitem__(self, idx):
        shp = self.shape
        return self._new(self.value[idx], [(self, lambda g: _scatter(shp, idx, g))])

    @property
    def T(self):
        return self._new(_T(self.value), [(self, lambda g: _T(g))])

    def flatten(self):
        shp = self.shape
        return self._new(self.value.flatten(), [(self, lambda g: g.reshape(shp))])

    def reshape(self, *shape):
        shp = self.shape
        return self._new(self.value.reshape(*shape), [(self, lambda g: g.reshape(shp))])

    def trace(self):
        n = self.shape[0]
        v = self.value
        return self._new(v.trace() if _is_dual(v) else np.trace(v), [(self, lambda g: g * np.eye(n))])

    # numpy function protocol: np.trace(box) -> box.trace()
    def __array_function__(self, func, types, args, kwargs):
        if func is np.trace:
            return args[0].trace()
        if func is np.transpose:
            return args[0].T
        return NotImplemented


def _backward(out, inputs):
    tape = out.tape
    grads = {id(out): 1.0}
    for node in reversed(tape.nodes):
        g = grads.pop(id(node), None)
        if g is None:
            continue
        if not node.parents:
            grads[id(node)] = g  # keep leaves
            continue
        for parent, vjp in node.parents:
            contrib = vjp(g)
            key = id(parent)
            if key in grads:
                grads[key] = grads[key] + contrib
            else:
                grads[key] = contrib
    res = []
    for x in inputs:
        g = grads.get(id(x))
        if g is None:
            g = np.zeros(_shape(x.value))
        res.append(g)
    return res


def gradient(function, args):
    """Euclidean gradient of function(*args) w.r.t. every argument (tuple of ndarrays)."""
    tape = _Tape()
    boxes = [Box(np.asarray(a, dtype=float), tape) for a in args]
    out = function(*boxes)
    if not isinstance(out, Box):  # constant function
        return tuple(np.zeros_like(np.asarray(a, dtype=float)) for a in args)
    grads = _backward(out, boxes)
    return tuple(np.asarray(_val(g), dtype=float) * np.ones(np.shape(a)) for g, a in zip(grads, args))


def hessian_vector_product(function, args, vectors):
    """d/dt grad f(args + t * vectors) at t = 0, per argument (tuple of ndarrays)."""
    tape = _Tape()
    boxes = [Box(Dual(a, v), tape) for a, v in zip(args, vectors)]
    out = function(*boxes)
    if not isinstance(out, Box):
        return tuple(np.zeros_like(np.asarray(a, dtype=float)) for a in args)
    grads = _backward(out, boxes)
    res = []
    for g, a in zip(grads, args):
        t = g.tan if _is_dual(g) else np.zeros(np.shape(a))
        res.append(np.asarray(t, dtype=float) * np.ones(np.shape(a)))
    return tuple(res)
