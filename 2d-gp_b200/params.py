"""Minimal stand-in for GPy's Param: a float with a name, a settable ``.gradient`` and a
constraint, so kernels keep the reference's attribute surface (myKernel.py:16-22) without
a GPy dependency."""
from __future__ import annotations

import math

import numpy as np

_LIM = 36.0


class Param:
    def __init__(self, name, value):
        self.name = name
        self.value = float(np.asarray(value).reshape(-1)[0])
        self.gradient = 0.0
        self.constraint = None          # None | 'positive' | ('bounded', lo, hi) | 'fixed'

    # --- GPy-like constraint declarations (myKernel.py:19-21) -----------------------------
    def constrain_positive(self):
        self.constraint = "positive"
        return self

    def constrain_bounded(self, lo, hi):
        self.constraint = ("bounded", float(lo), float(hi))
        return self

    def constrain_fixed(self, value=None):
        if value is not None:
            self.value = float(value)
        if self.constraint != "fixed":
            self._before_fix = self.constraint
        self.constraint = "fixed"
        return self

    fix = constrain_fixed

    def unconstrain_fixed(self):
        """GPy's unfix(): back to the constraint the parameter had before it was fixed."""
        if self.constraint == "fixed":
            self.constraint = getattr(self, "_before_fix", None)
        return self

    unfix = unconstrain_fixed

    # --- numeric behaviour ------------------------------------------------------------------
    def __float__(self):
        return self.value

    def __getitem__(self, i):            # GPy idiom ``self.var[0]`` (myKernel.py:352)
        return self.value

    def __array__(self, dtype=None, copy=None):
        return np.array([self.value], dtype=dtype or np.float64)

    @property
    def values(self):
        return np.array([self.value])

    def __repr__(self):
        return "Param(%s=%r)" % (self.name, self.value)

    def _b(self, other, op):
        return op(self.value, float(other))

    __add__ = lambda s, o: s.value + float(o)
    __radd__ = __add__
    __sub__ = lambda s, o: s.value - float(o)
    __rsub__ = lambda s, o: float(o) - s.value
    __mul__ = lambda s, o: s.value * float(o)
    __rmul__ = __mul__
    __truediv__ = lambda s, o: s.value / float(o)
    __rtruediv__ = lambda s, o: float(o) / s.value
    __pow__ = lambda s, o: s.value ** o
    __neg__ = lambda s: -s.value
    __lt__ = lambda s, o: s.value < float(o)
    __le__ = lambda s, o: s.value <= float(o)
    __gt__ = lambda s, o: s.value > float(o)
    __ge__ = lambda s, o: s.value >= float(o)

    # --- unconstrained <-> constrained (GPy Logexp / Logistic transforms) -------------------
    def to_free(self):
        v, c = self.value, self.constraint
        if c == "positive":
            return v if v > _LIM else math.log(math.expm1(v))          # inverse softplus
        if isinstance(c, tuple):
            lo, hi = c[1], c[2]
            p = min(max((v - lo) / (hi - lo), 1e-10), 1 - 1e-10)
            return math.log(p / (1.0 - p))
        return v

    def from_free(self, x):
        c = self.constraint
        if c == "positive":
            # GPy's Logexp clips its argument the same way: the value never underflows to 0
            x = max(x, -_LIM)
            self.value = x if x > _LIM else math.log1p(math.exp(x))
        elif isinstance(c, tuple):
            lo, hi = c[1], c[2]
            self.value = lo + (hi - lo) / (1.0 + math.exp(-min(max(x, -_LIM), _LIM)))
        elif c != "fixed":
            self.value = float(x)

    def dvalue_dfree(self, x):
        c = self.constraint
        if c == "positive":
            return 1.0 if x > _LIM else 1.0 / (1.0 + math.exp(-max(x, -_LIM)))
        if isinstance(c, tuple):
            lo, hi = c[1], c[2]
            s = 1.0 / (1.0 + math.exp(-min(max(x, -_LIM), _LIM)))
            return (hi - lo) * s * (1.0 - s)
        return 1.0
