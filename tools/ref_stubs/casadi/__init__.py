"""Stub of the casadi names the reference car env pulls in with `from casadi import *`.
Symbolic objects are inert; MX.tanh must return a numeric tanh for numeric input because the
reference's Euler step calls it on a float."""
import types  # noqa  (re-exported: the reference uses `types.SimpleNamespace` via the star import)
import numpy as np


class _Sym:
    shape = (1, 1)

    def _op(self, *a, **k):
        return _Sym()
    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = _op
    __truediv__ = __rtruediv__ = __pow__ = __rpow__ = __neg__ = _op


class _Vec(_Sym):
    def __init__(self, n):
        self.shape = (n, 1)


class MX(_Sym):
    @staticmethod
    def sym(name, *a):
        return _Sym()

    @staticmethod
    def tanh(x):
        if isinstance(x, _Sym):
            return _Sym()
        return np.tanh(x)

    @staticmethod
    def cos(x):
        return _Sym() if isinstance(x, _Sym) else np.cos(x)

    @staticmethod
    def sin(x):
        return _Sym() if isinstance(x, _Sym) else np.sin(x)


def vertcat(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        return _Vec(len(args[0]))
    return _Vec(len(args))


def Function(*a, **k):
    return None


__all__ = ["types", "MX", "vertcat", "Function", "np"]
