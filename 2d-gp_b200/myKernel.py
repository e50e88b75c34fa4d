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
