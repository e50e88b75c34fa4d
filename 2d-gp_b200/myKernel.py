"""Kernel plug-ins with the reference's names, constructor arguments and methods
(myKernel.py), backed by the CUDA covariance build.  No GPy needed: the classes expose the
subset of the ``Kern`` protocol GPy's GPRegression calls -- ``K(X, X2)``, ``Kdiag(X)``,
``update_gradients_full(dL_dK, X, X2)`` and parameters with ``.gradient``.

    myKernel(2, [0, 1], l_df, l_cf, ratio)   ratio * divergence-free + (1-ratio) * curl-free
    nonDivK(2, [0, 1], length)               divergence-free only   (ratio = 1)
    nonRotK(2, [0, 1], l)                    curl-free only         (ratio = 0)

``update_gradients_full`` defaults to the analytic derivative; set
``kernel.reference_compat = True`` to reproduce the reference's integrands
(myKernel.py:77-81, 91-96), which are not the derivative of its own K.
"""
from __future__ import annotations

import numpy as np

from . import engine
from .params import Param


class _HelmholtzBase:
    name = "helmholtz"
    reference_compat = False

    def __init__(self, input_dim, active_dim):
        assert input_dim == 2, "For this kernel we assume input_dim=2"       # myKernel.py:15
        self.input_dim = input_dim
        self.active_dims = list(active_dim)

    # theta = (l_df, l_cf, ratio) as floats
    def _theta(self):
        raise NotImplementedError

    def _slice(self, X):
        X = np.asarray(X) if not hasattr(X, "device") else X
        if X.shape[1] != 2:
            X = X[:, self.active_dims]
        return X

    def parameters_changed(self):
        pass

    def K_dev(self, X, X2=None):
        """Covariance as a device tensor [2N, 2M] (reference block layout)."""
        l_df, l_cf, ratio = self._theta()
        return engine.kernel_K(self._slice(X), None if X2 is None else self._slice(X2), l_df, l_cf, ratio)

    def K(self, X, X2=None):
        return self.K_dev(X, X2).cpu().numpy()

    def Kdiag(self, X):
        l_df, l_cf, ratio = self._theta()
        # the reference replicates the prior variance X.shape[0]*X.shape[1] times (myKernel.py:57)
        return engine.kernel_Kdiag(np.shape(X)[0] * np.shape(X)[1] // 2, l_df, l_cf, ratio).cpu().numpy()

    def _grad_sums(self, dL_dK, X, X2):
        l_df, l_cf, ratio = self._theta()
        g = engine.kernel_grad_sums(dL_dK, self._slice(X), None if X2 is None else self._slice(X2),
                                    l_df, l_cf, ratio, reference_compat=self.reference_compat)
        return g.cpu().numpy()

    def update_gradients_diag(self, dL_dKdiag, X):
        pass                                                                   # myKernel.py:108-109

    def gradients_X(self, dL_dK, X, X2=None):
        # the reference implementation raises before computing anything (myKernel.py:123)
        raise NotImplementedError("gradients_X is not part of the reference contract (it raises there)")

    def gradients_X_diag(self, dL_dKdiag, X):
        pass                                                                   # myKernel.py:144-146

    @property
    def param_array(self):
        return np.array([float(p) for p in self.parameters])

    def copy(self):
        import copy
        return copy.deepcopy(self)

    def __mul__(self, other):
        if isinstance(other, Kt):
            return SpaceTimeKern(other, self)
        return NotImplemented


class myKernel(_HelmholtzBase):
    """myKernel.myKernel (myKernel.py:12-146)."""

    def __init__(self, input_dim, active_dim=[0, 1], l_df=1., l_cf=1, ratio=1.):
        super().__init__(input_dim, active_dim)
        self.name = "myKern"
        self.length_df = Param("length_df", l_df).constrain_positive()
        self.length_cf = Param("length_cf", l_cf).constrain_positive()
        self.ratio = Param("ratio", ratio).constrain_bounded(0, 1)
        self.parameters = [self.length_df, self.length_cf, self.ratio]

    def _theta(self):
        return float(self.length_df), float(self.length_cf), float(self.ratio)

    def update_gradients_full(self, dL_dK, X, X2=None):
        g = self._grad_sums(dL_dK, X, X2)
        self.length_df.gradient, self.length_cf.gradient, self.ratio.gradient = map(float, g)


class nonDivK(_HelmholtzBase):
    """Divergence-free kernel, myKernel.nonDivK (myKernel.py:148-242)."""

    def __init__(self, input_dim, active_dim=[0, 1], length=1.):
        super().__init__(input_dim, active_dim)
        self.name = "nonDivK"
        self.length = Param("length", length).constrain_positive()
        self.parameters = [self.length]

    def _theta(self):
        return float(self.length), 1.0, 1.0

    def update_gradients_full(self, dL_dK, X, X2=None):
        self.length.gradient = float(self._grad_sums(dL_dK, X, X2)[0])


class nonRotK(_HelmholtzBase):
    """Curl-free kernel, myKernel.nonRotK (myKernel.py:244-334).  The reference stores its
    gradient on a non-existent ``length_cf`` (myKernel.py:301); here it lands on ``length``."""

    def __init__(self, input_dim, active_dim=[0, 1], l=1.):
        super().__init__(input_dim, active_dim)
        self.name = "nonRotK"
        self.length = Param("length", l).constrain_positive()
        self.parameters = [self.length]

    def _theta(self):
        return 1.0, float(self.length), 0.0

    def update_gradients_full(self, dL_dK, X, X2=None):
        self.length.gradient = float(self._grad_sums(dL_dK, X, X2)[1])


class Kt:
    """Time kernel of myKernel.Kt (myKernel.py:337-391): var * exp(-dt^2 / 2 lengthscale^2) tiled
    over the 2x2 velocity blocks.  Used as ``Kt(...) * nonDivK(...)`` (scratch.py:506-508).  The
    reference's gradient method is a stub (myKernel.py:365-372); here it is the derivative."""

    def __init__(self, input_dim=1, active_dims=[0], var=1, lengthscale=1.):
        assert input_dim == 1, "For this kernel we assume input_dim=1"          # myKernel.py:340
        self.input_dim, self.active_dims, self.name = 1, list(active_dims), "Kt"
        self.var = Param("var", var).constrain_positive()
        self.lengthscale = Param("lengthscale", lengthscale).constrain_positive()
        self.parameters = [self.var, self.lengthscale]

    def parameters_changed(self):
        pass

    def _t(self, X):
        X = np.asarray(X, dtype=np.float64)
        return X[:, self.active_dims] if X.shape[1] != 1 else X

    def K(self, X, X2=None):
        C = engine.rbf_K(self._t(X), None if X2 is None else self._t(X2), [float(self.var)], [[float(self.lengthscale)]])
        return C.repeat(2, 2).cpu().numpy()                                   # [[C, C], [C, C]]  (myKernel.py:357-358)

    def Kdiag(self, X):
        return np.ones(np.shape(X)[0] * np.shape(X)[1] * 2) * float(self.var)     # myKernel.py:362-363

    def update_gradients_full(self, dL_dK, X, X2=None):
        W = np.asarray(dL_dK, dtype=np.float64)
        n, m = W.shape[0] // 2, W.shape[1] // 2
        Wf = W[:n, :m] + W[:n, m:] + W[n:, :m] + W[n:, m:]
        g = engine.rbf_grad_sums(Wf, self._t(X), None if X2 is None else self._t(X2), [float(self.var)],
                                 [[float(self.lengthscale)]]).cpu().numpy()
        self.var.gradient, self.lengthscale.gradient = float(g[0]), float(g[1])

    def update_gradients_diag(self, dL_dKdiag, X):
        pass

    def gradients_X_diag(self, dL_dKdiag, X):
        pass

    def copy(self):
        import copy
        return copy.deepcopy(self)

    def __mul__(self, other):
        if isinstance(other, _HelmholtzBase):
            return SpaceTimeKern(self, other)
        return NotImplemented


class SpaceTimeKern:
    """``Kt(t) * <Helmholtz kernel>(a, b)``: GPy's product of kernels acting on different columns
    (scratch.py:506-508).  Parameters in GPy's order: the time part first, then the space part."""

    family = "spacetime"

    def __init__(self, kt, kxy):
        self.kt, self.kxy = kt.copy(), kxy.copy()
        self.name = "mul"

    @property
    def parameters(self):
        return list(self.kt.parameters) + list(self.kxy.parameters)

    @property
    def param_array(self):
        return np.array([float(p) for p in self.parameters])

    def parameter_names(self):
        return ["mul.Kt." + p.name for p in self.kt.parameters] + \
               ["mul.%s.%s" % (self.kxy.name, p.name) for p in self.kxy.parameters]

    def theta5(self):
        """(l_df, l_cf, ratio, tvar, lt)."""
        return tuple(self.kxy._theta()) + (float(self.kt.var), float(self.kt.lengthscale))

    def points3(self, X):
        X = np.asarray(X, dtype=np.float64)
        return np.ascontiguousarray(X[:, list(self.kt.active_dims) + list(self.kxy.active_dims)])

    def K(self, X, X2=None):
        return engine.st_K(self.points3(X), None if X2 is None else self.points3(X2), *self.theta5()).cpu().numpy()

    def Kdiag(self, X):
        l_df, l_cf, ratio, tvar, _ = self.theta5()
        return np.full(2 * np.shape(X)[0], tvar * (ratio / l_df ** 2 + (1 - ratio) / l_cf ** 2))

    def scatter_gradient(self, g5):
        """g5 ordered (l_df, l_cf, ratio, tvar, lt), the engine's order."""
        self.kt.var.gradient, self.kt.lengthscale.gradient = float(g5[3]), float(g5[4])
        if isinstance(self.kxy, myKernel):
            self.kxy.length_df.gradient, self.kxy.length_cf.gradient, self.kxy.ratio.gradient = map(float, g5[:3])
        elif isinstance(self.kxy, nonDivK):
            self.kxy.length.gradient = float(g5[0])
        else:
            self.kxy.length.gradient = float(g5[1])

    def update_gradients_full(self, dL_dK, X, X2=None):
        g = engine.st_grad_sums(dL_dK, self.points3(X), None if X2 is None else self.points3(X2), *self.theta5())
        self.scatter_gradient(g.cpu().numpy())

    def copy(self):
        import copy
        return copy.deepcopy(self)
