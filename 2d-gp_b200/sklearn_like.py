"""The slice of scikit-learn's GP API that krig.scikit_prior uses (krig.py:5,174-194), on the CUDA
engine:

    from gp2d_b200.sklearn_like import GaussianProcessRegressor, kernels
    k = HP[0] * kernels.RBF(length_scale=[lt, ly, lx]) + kernels.WhiteKernel(noise_level=noise)
    m = GaussianProcessRegressor(kernel=k, optimizer=None).fit(XT, u)
    U, Ustd = m.predict(X, return_std=True)

Same conventions as scikit-learn: ``alpha`` (1e-10) is added to the diagonal, the WhiteKernel is
part of the predictive variance, ``log_marginal_likelihood_value_`` follows _gpr.py:613-615.
``optimizer=None`` (krig.py:182) keeps the hyper-parameters fixed; the default
``optimizer="fmin_l_bfgs_b"`` (testKrig.py:139-140,151-152) maximises the log-marginal likelihood
over log(theta) inside the kernels' bounds with ``n_restarts_optimizer`` extra log-uniform starts
(_gpr.py:296-333) -- every evaluation is one gp2d_rbf_lml_grad call on the GPU."""
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
        self.length_scale_bounds = length_scale_bounds

    def __repr__(self):
        return "RBF(length_scale=%s)" % (np.round(np.atleast_1d(self.length_scale), 3).tolist(),)


class WhiteKernel(_Kernel):
    def __init__(self, noise_level=1.0, noise_level_bounds=(1e-5, 1e5)):
        self.noise_level = float(noise_level)
        self.noise_level_bounds = noise_level_bounds

    def __repr__(self):
        return "WhiteKernel(noise_level=%.3g)" % self.noise_level


class ConstantKernel(_Kernel):
    def __init__(self, constant_value=1.0, constant_value_bounds=(1e-5, 1e5)):
        self.constant_value = float(constant_value)
        self.constant_value_bounds = constant_value_bounds

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


def _hyperparameters(k, D, out):
    """Free hyper-parameters of the expression in scikit-learn's order (depth first, k1 before k2):
    list of (object, attribute, index or None, bounds)."""
    if isinstance(k, (Sum, Product)):
        _hyperparameters(k.k1, D, out)
        _hyperparameters(k.k2, D, out)
    elif isinstance(k, ConstantKernel):
        if k.constant_value_bounds != "fixed":
            out.append((k, "constant_value", None, k.constant_value_bounds))
    elif isinstance(k, WhiteKernel):
        if k.noise_level_bounds != "fixed":
            out.append((k, "noise_level", None, k.noise_level_bounds))
    elif isinstance(k, RBF):
        if k.length_scale_bounds != "fixed":
            n = np.size(k.length_scale)
            if n == 1:
                out.append((k, "length_scale", None, k.length_scale_bounds))
            else:
                k.length_scale = np.array(k.length_scale, dtype=np.float64)
                out += [(k, "length_scale", i, k.length_scale_bounds) for i in range(n)]
    return out


def _get(h):
    obj, attr, i, _ = h
    v = getattr(obj, attr)
    return float(np.asarray(v).reshape(-1)[0] if i is None else v[i])


def _set(h, value):
    obj, attr, i, _ = h
    if i is None:
        setattr(obj, attr, float(value))
    else:
        getattr(obj, attr)[i] = float(value)


class GaussianProcessRegressor:
    def __init__(self, kernel=None, alpha=1e-10, optimizer="fmin_l_bfgs_b", n_restarts_optimizer=0, normalize_y=False,
                 copy_X_train=True, random_state=None):
        if optimizer not in (None, "fmin_l_bfgs_b"):
            raise NotImplementedError("optimizer must be None or 'fmin_l_bfgs_b'")
        if normalize_y:
            raise NotImplementedError("normalize_y is not used by the reference")
        self.kernel = kernel if kernel is not None else ConstantKernel(1.0, "fixed") * RBF(1.0, "fixed")
        self.alpha = float(alpha)
        self.optimizer = optimizer
        self.n_restarts_optimizer = int(n_restarts_optimizer)
        self.random_state = random_state

    # theta (log of the free hyper-parameters) of the fitted kernel, like sklearn's kernel_.theta
    def _theta_of(self, k, D):
        return np.log([_get(h) for h in _hyperparameters(k, D, [])])

    def _neg_lml_and_grad(self, theta, hyper, D):
        """Objective of the search: -LML and its gradient w.r.t. log(theta) (chain rule theta * d/dtheta)."""
        import copy
        for h, t in zip(hyper, theta):
            _set(h, np.exp(t))
        var, ls, noise = _flatten(self.kernel_, D)
        self._gp.set_params(var, ls, noise)
        try:
            lml, g = self._gp.lml_and_grad()        # (var_q, ls_q[0..D-1])_q, noise
        except np.linalg.LinAlgError:
            return 1e25, np.zeros(len(theta))
        # map the engine's gradient onto the free hyper-parameters
        grad = np.zeros(len(theta))
        comps = self._components                   # per RBF component: (constants, rbf)
        for i, h in enumerate(hyper):
            obj, attr, idx, _ = h
            if attr == "noise_level":
                grad[i] = g[-1] * _get(h)
                continue
            for q, (consts, rbf) in enumerate(comps):
                blk = g[q * (1 + D):(q + 1) * (1 + D)]
                if attr == "constant_value" and any(obj is c for c in consts):
                    grad[i] += blk[0] * var[q]                       # d/dlog c = var_q * d/dvar_q
                elif attr == "length_scale" and obj is rbf:
                    grad[i] += (blk[1 + idx] * ls[q][idx]) if idx is not None else float(np.sum(blk[1:] * ls[q]))
        return -lml, -grad

    def fit(self, X, y):
        import copy
        import scipy.optimize as sopt
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        self._y_2d = y.ndim == 2
        D = X.shape[1]
        self.kernel_ = copy.deepcopy(self.kernel)
        var, ls, noise = _flatten(self.kernel_, D)
        if not var:
            raise ValueError("the kernel has no RBF term")
        self._components = _components(self.kernel_)
        self._gp = ScalarGP(X, y.reshape(-1), var, ls, noise, jitter=self.alpha)
        hyper = _hyperparameters(self.kernel_, D, [])
        if self.optimizer is not None and hyper:
            bounds = [tuple(np.log(h[3])) for h in hyper]
            starts = [np.log([_get(h) for h in hyper])]
            rng = np.random.RandomState(self.random_state)
            for _ in range(self.n_restarts_optimizer):
                starts.append(np.array([rng.uniform(lo, hi) for lo, hi in bounds]))
            best = None
            for x0 in starts:
                res = sopt.minimize(self._neg_lml_and_grad, x0, args=(hyper, D), jac=True, method="L-BFGS-B", bounds=bounds)
                if best is None or res.fun < best.fun:
                    best = res
            for h, t in zip(hyper, best.x):
                _set(h, np.exp(t))
            var, ls, noise = _flatten(self.kernel_, D)
            self._gp.set_params(var, ls, noise)
        self._noise = noise
        self.log_marginal_likelihood_value_ = self._gp.fit()
        self.X_train_, self.y_train_ = X, y
        return self

    # pickling (krig.scikitSnapshot loads <name>_scikit_u.pkl, krig.py:237-242): the device state is not
    # stored, only the fitted kernel and the training data; loading refits with fixed hyper-parameters
    def __getstate__(self):
        st = {k: v for k, v in self.__dict__.items() if k not in ("_gp", "_components")}
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        if "kernel_" in st:
            D = self.X_train_.shape[1]
            var, ls, noise = _flatten(self.kernel_, D)
            self._components = _components(self.kernel_)
            self._gp = ScalarGP(self.X_train_, np.asarray(self.y_train_).reshape(-1), var, ls, noise, jitter=self.alpha)
            self._gp.fit()

    @property
    def theta_(self):
        return self._theta_of(self.kernel_, self.X_train_.shape[1])

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


def _components(k):
    """Per RBF term of the sum: ([constant kernels multiplying it], rbf), in _flatten's order."""
    if isinstance(k, Sum):
        return _components(k.k1) + _components(k.k2)
    if isinstance(k, WhiteKernel):
        return []
    consts, node = [], k
    while isinstance(node, Product):
        if isinstance(node.k1, ConstantKernel):
            consts.append(node.k1)
            node = node.k2
        else:
            consts.append(node.k2)
            node = node.k1
    return [(consts, node)]
