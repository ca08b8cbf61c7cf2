"""Shim of pymanopt.tools (only what src/solver/RIPTRM.py:12 imports)."""
import functools


class ndarraySequenceMixin:
    __array_priority__ = 1000
    __array_ufunc__ = None


def return_as_class_instance(method=None, *, unpack=True):
    def make_wrapper(function):
        @functools.wraps(function)
        def wrapper(self, *args, **kwargs):
            return_value = function(self, *args, **kwargs)
            if unpack:
                return self.__class__(*return_value)
            return self.__class__(return_value)
        return wrapper
    if method is not None and callable(method):
        return make_wrapper(method)
    return make_wrapper
