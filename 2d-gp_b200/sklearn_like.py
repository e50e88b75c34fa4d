"""The slice of scikit-learn's GP API that krig.scikit_prior uses (krig.py:5,174-194), on the CUDA
engine:

    from gp2d_b200.sklearn_like import GaussianProcessRegressor, kernels
    k = HP[0] * kernels.RBF(length_scale=[lt, ly, lx]) + kernels.WhiteKernel(noise_level=noise)
    m = GaussianProcessRegressor(kernel=k, optimizer=None).fit(XT, u)
    U, Ustd = m.predict(X, return_std=True)

Same conventions as scikit-learn: ``alpha`` (1e-10) is added to the diagonal, the WhiteKernel is
part of the predictive variance, ``log_marginal_likelihood_value_`` follows _gpr.py:613-615.
Only fixed hyper-parameters (``optimizer=None``), which is how the reference calls it; the
hyper-parameter search is models.GPRegression.optimize_restarts."""
from __future__ import annotations

import numpy as np

from .engine import ScalarGP


class _Kernel:
    def __add__(self, other):
        return Sum(self, other)

    def __radd__(self, other):
        return Sum(other, self) if isinstance(other, _Kernel) else NotImplemented

    def __mul__(self, other):
        return Product(self, other if isinstance(other, _Kernel) else ConstantKernel(other))

    def __rmul__(self, other):
        return Product(other if isinstance(other, _Kernel) else ConstantKernel(other), self)


class RBF(_Kernel):
    def __init__(self, length_scale=1.0, length_scale_bounds=(1e-5, 1e5)):
        self.length_scale = length_scale

    def __repr__(self):
        return "RBF(length_scale=%s)" % (np.round(np.atleast_1d(self.length_scale), 3).tolist(),)


class WhiteKernel(_Kernel):
    def __init__(self, noise_level=1.0, noise_level_bounds=(1e-5, 1e5)):
        self.noise_level = float(noise_level)

    def __repr__(self):
        return "WhiteKernel(noise_level=%.3g)" % self.noise_level


class ConstantKernel(_Kernel):
    def __init__(self, constant_value=1.0, constant_value_bounds=(1e-5, 1e5)):
        self.constant_value = float(constant_value)

    def __repr__(self):
        return "%.3g**2" % np.sqrt(self.constant_value)


class Sum(_Kernel):
    def __init__(self, k1, k2):
        self.k1, self.k2 = k1, k2

    def __repr__(self):
        return "%r + %r" % (self.k1, self.k2)


class Product(_Kernel):
    def __init__(self, k1, k2):
        self.k1, self.k2 = k1, k2

    def __repr__(self):
        return "%r * %r" % (self.k1, self.k2)


class kernels:                       # ``from sklearn.gaussian_process import kernels`` look-alike
    RBF, WhiteKernel, ConstantKernel, Sum, Product = RBF, WhiteKernel, ConstantKernel, Sum, Product


def _flatten(k, D):
    """Kernel expression -> ([variance_q], [lengthscales_q[D]], white noise).  Supported terms of
    the sum: RBF, constant * RBF (any nesting of constants), WhiteKernel."""
    if isinstance(k, Sum):
        a, b = _flatten(k.k1, D), _flatten(k.k2, D)
        return a[0] + b[0], a[1] + b[1], a[2] + b[2]
    if isinstance(k, WhiteKernel):
        return [], [], k.noise_level
    c, node = 1.0, k
    while isinstance(node, Product):
        if isinstance(node.k1, ConstantKernel):
            c, node = c * node.k1.constant_value, node.k2
        elif isinstance(node.k2, ConstantKernel):
            c, node = c * node.k2.constant_value, node.k1
        else:
            raise NotImplementedError("only constant * RBF products are on the GPU path")
    if not isinstance(node, RBF):
        raise NotImplementedError("unsupported kernel term %r" % (node,))
    ls = np.atleast_1d(np.asarray(node.length_scale, dtype=np.float64))
    if ls.size == 1:
        ls = np.repeat(ls, D)
    if ls.size != D:
        raise ValueError("anisotropic length_scale must have %d entries" % D)
    return [c], [ls], 0.0


class GaussianProcessRegressor:
    def __init__(self, kernel=None, alpha=1e-10, optimizer=None, n_restarts_optimizer=0, normalize_y=False,
                 copy_X_train=True, random_state=None):
        if optimizer is not None:
            raise NotImplementedError("optimizer must be None (krig.py:182): fixed hyper-parameters; "
                                      "use models.GPRegression.optimize_restarts for the search")
        if normalize_y:
            raise NotImplementedError("normalize_y is not used by the reference")
        self.kernel = kernel if kernel is not None else ConstantKernel(1.0) * RBF(1.0)
        self.alpha = float(alpha)

    def fit(self, X, y):
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        self._y_2d = y.ndim == 2
        var, ls, noise = _flatten(self.kernel, X.shape[1])
        if not var:
            raise ValueError("the kernel has no RBF term")
        self.kernel_ = self.kernel
        self._noise = noise
        self._gp = ScalarGP(X, y.reshape(-1), var, ls, noise, jitter=self.alpha)
        self.log_marginal_likelihood_value_ = self._gp.fit()
        self.X_train_, self.y_train_ = X, y
        return self

    def log_marginal_likelihood(self, theta=None):
        if theta is not None:
            raise NotImplementedError("evaluate other hyper-parameters through models.GPRegression")
        return self.log_marginal_likelihood_value_

    def predict(self, X, return_std=False, return_cov=False):
        if return_cov:
            raise NotImplementedError("only marginal standard deviations (krig.py:194)")
        mean, var = self._gp.predict(np.asarray(X, dtype=np.float64), include_noise=True)
        mean = mean.cpu().numpy()
        if self._y_2d:
            mean = mean[:, None]
        if return_std:
            return mean, np.sqrt(var.cpu().numpy())
        return mean
