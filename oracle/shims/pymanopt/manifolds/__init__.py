"""Shim of pymanopt.manifolds: re-exports the restated formulas of oracle/manifolds.py."""
import sys
import types

from oracle import manifolds as _m

Sphere = _m.Sphere
Grassmann = _m.Grassmann
Stiefel = _m.Stiefel
Oblique = _m.Oblique
SkewSymmetric = _m.SkewSymmetric
SymmetricPositiveDefinite = _m.SymmetricPositiveDefinite
Product = _m.Product


def _submodule(name, **attrs):
    mod = types.ModuleType(f"{__name__}.{name}")
    mod.__dict__.update(attrs)
    sys.modules[mod.__name__] = mod
    return mod


sphere = _submodule("sphere", Sphere=Sphere)
grassmann = _submodule("grassmann", Grassmann=Grassmann)
stiefel = _submodule("stiefel", Stiefel=Stiefel)
oblique = _submodule("oblique", Oblique=Oblique)
euclidean = _submodule("euclidean", SkewSymmetric=SkewSymmetric)
positive_definite = _submodule("positive_definite", SymmetricPositiveDefinite=SymmetricPositiveDefinite)
product = _submodule("product", Product=Product, _ProductTangentVector=_m._TangentList)
