"""GPy-style scalar kernels backed by the CUDA covariance build: ``RBF`` (optionally ARD) and sums
of them (``k + k``), the kernels krig.kriging builds for kernelType=1
(``GPy.kern.RBF(input_dim=3, ARD=True)``, ``k = k + k2``; krig.py:388,405-407).

Parameter order follows GPy's ``param_array``: per part ``[variance, lengthscale...]`` -- the
layout krig.scikit_prior decodes (``HP[0]``, ``HP[1:4]``, ``HP[4]``, ``HP[5:8]``; krig.py:174-178).
"""
from __future__ import annotations

import numpy as np

from . import engine
from .params import Param


class _ScalarKern:
    family = "rbf"

    def parts_list(self):
        raise NotImplementedError

    # ---- what the engine needs ------------------------------------------------------------
    def rbf_params(self):
        """(variances[Q], lengthscales[Q][D]) of the sum."""
        parts = self.parts_list()
        return [float(p.variance) for p in parts], [p._ls_vector() for p in parts]

    def _slice(self, X):
        X = np.asarray(X, dtype=np.float64)
        dims = self.parts_list()[0].active_dims
        return X if X.shape[1] == len(dims) and dims == list(range(X.shape[1])) else X[:, dims]

    @property
    def parameters(self):
        out = []
        for p in self.parts_list():
            out += [p.variance] + p.lengthscale
        return out

    @property
    def param_array(self):
        return np.array([float(p) for p in self.parameters])

    def parameter_names(self):
        names = []
        for i, p in enumerate(self.parts_list()):
            base = p.name if i == 0 else "%s_%d" % (p.name, i)
            names.append(base + ".variance")
            names += [base + ".lengthscale"] if len(p.lengthscale) == 1 else \
                ["%s.lengthscale[%d]" % (base, d) for d in range(len(p.lengthscale))]
        return names

    # ---- Kern protocol -----------------------------------------------------------------------
    def K(self, X, X2=None):
        var, ls = self.rbf_params()
        return engine.rbf_K(self._slice(X), None if X2 is None else self._slice(X2), var, ls).cpu().numpy()

    def Kdiag(self, X):
        return np.full(np.shape(X)[0], sum(self.rbf_params()[0]))

    def update_gradients_full(self, dL_dK, X, X2=None):
        var, ls = self.rbf_params()
        g = engine.rbf_grad_sums(dL_dK, self._slice(X), None if X2 is None else self._slice(X2), var, ls).cpu().numpy()
        self._scatter_gradient(g)

    def _scatter_gradient(self, g):
        """g is ordered (variance_q, lengthscale_q[0..D-1]) per part (the engine's order)."""
        D = self.parts_list()[0].input_dim
        for q, p in enumerate(self.parts_list()):
            blk = g[q * (1 + D):(q + 1) * (1 + D)]
            p.variance.gradient = float(blk[0])
            if p.ARD:
                for d in range(D):
                    p.lengthscale[d].gradient = float(blk[1 + d])
            else:
                p.lengthscale[0].gradient = float(np.sum(blk[1:]))

    def update_gradients_diag(self, dL_dKdiag, X):
        for p in self.parts_list():
            p.variance.gradient = float(np.sum(dL_dKdiag))

    def copy(self):
        import copy
        return copy.deepcopy(self)

    def __add__(self, other):
        if not isinstance(other, _ScalarKern):
            return NotImplemented
        return Add(self.parts_list() + other.parts_list())

    def __mul__(self, other):
        if isinstance(self, RBF) and isinstance(other, RBF):
            return Prod([self, other])
        if isinstance(self, Prod) and isinstance(other, RBF):
            return Prod(self.factors + [other])
        return NotImplemented

    def parameters_changed(self):
        pass


class RBF(_ScalarKern):
    """variance * exp(-1/2 sum_d ((x_d - x'_d)/lengthscale_d)^2)   (GPy.kern.RBF)."""

    def __init__(self, input_dim, variance=1., lengthscale=None, ARD=False, active_dims=None, name="rbf"):
        if not 1 <= input_dim <= 4:
            raise ValueError("the GPU kernel supports 1 <= input_dim <= 4")
        self.input_dim, self.ARD, self.name = int(input_dim), bool(ARD), name
        self.active_dims = list(range(input_dim)) if active_dims is None else list(active_dims)
        self.variance = Param("variance", variance).constrain_positive()
        if lengthscale is None:
            lengthscale = np.ones(input_dim if ARD else 1)
        ls = np.atleast_1d(np.asarray(lengthscale, dtype=np.float64))
        if ARD and ls.size == 1:
            ls = np.repeat(ls, input_dim)
        if ls.size != (input_dim if ARD else 1):
            raise ValueError("lengthscale must have %d entries" % (input_dim if ARD else 1))
        self.lengthscale = [Param("lengthscale", v).constrain_positive() for v in ls]

    def parts_list(self):
        return [self]

    def _ls_vector(self):
        v = [float(p) for p in self.lengthscale]
        return v if self.ARD else v * self.input_dim


class Add(_ScalarKern):
    """Sum of RBF parts (GPy's ``k1 + k2``); every part acts on the same input columns."""

    def __init__(self, parts, name="sum"):
        parts = list(parts)
        if len(parts) > 4:
            raise ValueError("at most 4 RBF components")
        if any(p.input_dim != parts[0].input_dim or p.active_dims != parts[0].active_dims for p in parts):
            raise ValueError("all parts must share input_dim / active_dims")
        # GPy copies the parts when adding, so k + k gives independent parameters
        self.parts = [p.copy() for p in parts]
        self.name = name
        self.input_dim = parts[0].input_dim
        self.active_dims = list(parts[0].active_dims)

    def parts_list(self):
        return self.parts


class Prod(_ScalarKern):
    """Product of RBF kernels acting on disjoint input columns, ``k1 * k2`` in GPy
    (GP_plots.py:738-740: ``RBF(1, active_dims=[0]) * RBF(1, active_dims=[1])``).  Such a product is
    one anisotropic RBF over the union of the columns with variance prod(variance_i); GPy keeps
    every factor's variance as its own (redundant) parameter, and so does this class."""

    def __init__(self, factors, name="mul"):
        factors = [f.copy() for f in factors]
        dims = [d for f in factors for d in f.active_dims]
        if len(set(dims)) != len(dims):
            raise NotImplementedError("only products of RBFs on disjoint input columns are on the GPU path")
        if len(dims) > 4:
            raise ValueError("at most 4 input dimensions")
        self.factors, self.name = factors, name
        self.active_dims = dims
        self.input_dim = len(dims)
        self.ARD = True

    def parts_list(self):
        return [self]

    # the engine sees ONE anisotropic component over active_dims
    def rbf_params(self):
        return [float(np.prod([float(f.variance) for f in self.factors]))], \
               [[l for f in self.factors for l in f._ls_vector()]]

    def _slice(self, X):
        X = np.asarray(X, dtype=np.float64)
        return X if X.shape[1] == len(self.active_dims) and self.active_dims == list(range(X.shape[1])) \
            else X[:, self.active_dims]

    @property
    def parameters(self):
        out = []
        for f in self.factors:
            out += [f.variance] + f.lengthscale
        return out

    def parameter_names(self):
        names = []
        for i, f in enumerate(self.factors):
            base = "mul.%s" % (f.name if i == 0 else "%s_%d" % (f.name, i))
            names.append(base + ".variance")
            names += [base + ".lengthscale"] if len(f.lengthscale) == 1 else \
                ["%s.lengthscale[%d]" % (base, d) for d in range(len(f.lengthscale))]
        return names

    def _scatter_gradient(self, g):
        """g = (d/dvariance, d/dl_0 .. d/dl_{D-1}) of the combined component."""
        total = float(np.prod([float(f.variance) for f in self.factors]))
        o = 1
        for f in self.factors:
            f.variance.gradient = float(g[0]) * total / float(f.variance)
            if f.ARD:
                for d in range(f.input_dim):
                    f.lengthscale[d].gradient = float(g[o + d])
            else:
                f.lengthscale[0].gradient = float(np.sum(g[o:o + f.input_dim]))
            o += f.input_dim
