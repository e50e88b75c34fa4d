"""The module krig.kriging imports for kernelType 2, 3 and 4 (krig.py:396-404) and which is
missing from the reference repository:

    k2 = myKernel2.divFreeK(input_dim=3, active_dims=[0,1,2], var=1., lt=1., ly=1., lx=1.)
    k2 = myKernel2.curlFreeK(input_dim=3, active_dims=[0,1,2], var=1., lt=1., ly=1., lx=1.)
    k2 = myKernel2.divFreeK(input_dim=3) + myKernel2.curlFreeK(input_dim=3)
    k = k2.copy();  k = k + k2  (nKernels - 1 times, krig.py:405-407)

Constructor names and arguments are the ones of those call sites; the covariance is the
divergence-free / curl-free construction of myKernel.py:39-52 applied to an anisotropic space-time
squared exponential var exp(-dt^2/2lt^2 - dy^2/2ly^2 - dx^2/2lx^2) (csrc/hsum.cuh).  Inputs are
rows (t, y, x) (krig.py "Pay attention on the order T,Y,X"), observations stacked [v; u]
(krig.py:392): the first block belongs to the first spatial coordinate.  input_dim=2 drops the
time factor (rows (y, x)).  Up to 8 terms per sum; every term keeps its own parameters, like GPy.
"""
from __future__ import annotations

import numpy as np

from . import engine
from .params import Param

MAX_TERMS = engine.HSUM_MAXQ


class _HelmholtzSumKern:
    """Protocol shared by a single term and a sum of terms: ``terms_list()`` and the GPy ``Kern``
    subset GPRegression needs."""

    family = "hsum"

    def terms_list(self):
        raise NotImplementedError

    @property
    def parameters(self):
        return [p for t in self.terms_list() for p in t._own_parameters()]

    @property
    def param_array(self):
        return np.array([float(p) for p in self.parameters])

    def parameter_names(self):
        terms = self.terms_list()
        if len(terms) == 1 and terms[0] is self:
            return ["%s.%s" % (self.name, p.name) for p in self._own_parameters()]
        seen, names = {}, []
        for t in terms:                         # GPy renames duplicates: divFreeK, divFreeK_1, ...
            i = seen.get(t.name, 0)
            seen[t.name] = i + 1
            base = "sum.%s" % (t.name if i == 0 else "%s_%d" % (t.name, i))
            names += ["%s.%s" % (base, p.name) for p in t._own_parameters()]
        return names

    def hsum_params(self):
        """(types[Q], params[Q,4] = var, lt, la, lb): what the engine takes."""
        terms = self.terms_list()
        return [t.TYPE for t in terms], [[float(t.var), float(t.lt), float(t.ly), float(t.lx)] for t in terms]

    def points(self, X):
        X = np.asarray(X, dtype=np.float64)
        d = self.terms_list()[0].active_dims
        if X.shape[1] == len(d) and d == list(range(len(d))):
            return np.ascontiguousarray(X)
        return np.ascontiguousarray(X[:, d])

    def K(self, X, X2=None):
        ty, pr = self.hsum_params()
        return engine.hsum_K(self.points(X), None if X2 is None else self.points(X2), ty, pr).cpu().numpy()

    def Kdiag(self, X):
        ty, pr = self.hsum_params()
        return engine.hsum_Kdiag(int(np.shape(X)[0]), self.terms_list()[0].input_dim, ty, pr).cpu().numpy()

    def scatter_gradient(self, g):
        """g[Q,4] (or flat) ordered (var, lt, la, lb) per term, the engine's order."""
        g = np.asarray(g, dtype=np.float64).reshape(-1, 4)
        for t, row in zip(self.terms_list(), g):
            t.var.gradient, t.lt.gradient, t.ly.gradient, t.lx.gradient = map(float, row)

    def update_gradients_full(self, dL_dK, X, X2=None):
        ty, pr = self.hsum_params()
        g = engine.hsum_grad_sums(dL_dK, self.points(X), None if X2 is None else self.points(X2), ty, pr)
        self.scatter_gradient(g.cpu().numpy())

    def update_gradients_diag(self, dL_dKdiag, X):
        pass

    def gradients_X(self, dL_dK, X, X2=None):
        raise NotImplementedError("gradients_X is not part of the reference contract (it raises there, myKernel.py:123)")

    def gradients_X_diag(self, dL_dKdiag, X):
        pass

    def parameters_changed(self):
        pass

    def copy(self):
        import copy
        return copy.deepcopy(self)

    def __add__(self, other):
        if not isinstance(other, _HelmholtzSumKern):
            return NotImplemented
        return HelmholtzSum(self.terms_list() + other.terms_list())


class _Term(_HelmholtzSumKern):
    TYPE = 0

    def __init__(self, input_dim=3, active_dims=None, var=1., lt=1., ly=1., lx=1.):
        if input_dim not in (2, 3):
            raise ValueError("input_dim is 3 (t, y, x) or 2 (y, x)")
        self.input_dim = int(input_dim)
        self.active_dims = list(range(input_dim)) if active_dims is None else list(active_dims)
        if len(self.active_dims) != self.input_dim:
            raise ValueError("active_dims must list input_dim columns")
        self.var = Param("var", var).constrain_positive()
        self.lt = Param("lt", lt).constrain_positive()
        self.ly = Param("ly", ly).constrain_positive()
        self.lx = Param("lx", lx).constrain_positive()

    def _own_parameters(self):
        # without a time column lt does not enter the covariance and is not a model parameter
        return [self.var, self.lt, self.ly, self.lx] if self.input_dim == 3 else [self.var, self.ly, self.lx]

    def terms_list(self):
        return [self]


class divFreeK(_Term):
    """Divergence-free term (krig.py:397)."""
    TYPE = 0
    name = "divFreeK"


class curlFreeK(_Term):
    """Curl-free term (krig.py:401)."""
    TYPE = 1
    name = "curlFreeK"


class HelmholtzSum(_HelmholtzSumKern):
    """``k1 + k2 + ...`` of divFreeK / curlFreeK terms acting on the same input columns."""

    name = "sum"

    def __init__(self, terms):
        terms = list(terms)
        if not 1 <= len(terms) <= MAX_TERMS:
            raise ValueError("between 1 and %d terms" % MAX_TERMS)
        if any(t.input_dim != terms[0].input_dim or t.active_dims != terms[0].active_dims for t in terms):
            raise ValueError("all terms must share input_dim / active_dims")
        self.terms = [t.copy() for t in terms]       # GPy copies the parts: k + k has independent parameters
        self.input_dim = terms[0].input_dim
        self.active_dims = list(terms[0].active_dims)

    def terms_list(self):
        return self.terms
