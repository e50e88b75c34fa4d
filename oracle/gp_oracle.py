"""fp64 numpy/scipy restatement of the reference's GP hot path.  TEST INFRASTRUCTURE.

Every function cites the reference lines it restates (paths relative to the
upstream repo rafaelcgon/2D-GP).  Nothing here is imported by the product
package; see ``oracle/__init__.py`` for who may call it and for the pinning
status of each part.

Conventions:
  * X is [N,2] (x, y) in km; y is the stacked observation vector [u; v] of
    length 2N (GP_laser.py:98,174; GP_plots.py:721-722).
  * Covariances use the reference's component-major block layout: row c*N+i,
    column c'*M+j (myKernel.py:40-43; GP_scripts.py:89-95).
  * theta = (l_df, l_cf, ratio); noise is the Gaussian noise *variance*.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla

__all__ = [
    "helmholtz_K", "helmholtz_Kdiag", "helmholtz_dK", "kernel_grad_sums",
    "fit", "predict", "lml", "lml_and_grad", "fit_predict_inverse_form",
    "haversine_km", "simlaser_inputs", "rbf_ard_K", "LOG_2PI",
    "kt_K", "st_K", "st_dK", "st_fit", "st_predict", "st_lml_and_grad",
    "rbf_sum_K", "rbf_sum_dK", "rbf_kernel_grad_sums", "rbf_fit", "rbf_predict", "rbf_lml_and_grad",
    "hsum_K", "hsum_Kdiag", "hsum_dK", "hsum_kernel_grad_sums", "hsum_fit", "hsum_predict", "hsum_lml_and_grad",
]

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# kernel
# --------------------------------------------------------------------------------------
def _pair_terms(X, X2):
    """dx1, dx2 and the outer-product terms (myKernel.py:30-35)."""
    X = np.asarray(X, dtype=np.float64)
    X2 = X if X2 is None else np.asarray(X2, dtype=np.float64)
    dx1 = X[:, 0][:, None] - X2[:, 0]
    dx2 = X[:, 1][:, None] - X2[:, 1]
    return dx1, dx2, dx1 * dx1, dx1 * dx2, dx2 * dx2


def _blk(a11, a12, a21, a22):
    """2x2 block stacking used throughout the reference (myKernel.py:40-41)."""
    return np.concatenate([np.concatenate([a11, a12], axis=1),
                           np.concatenate([a21, a22], axis=1)], axis=0)


def helmholtz_K(X, X2, l_df, l_cf, ratio):
    """ratio*K_divfree + (1-ratio)*K_curlfree, [2N,2M].

    Restates myKernel.myKernel.K (myKernel.py:27-53), identical to
    GP_scripts.myKernel (GP_scripts.py:6-42).  ratio=1 gives nonDivK.K
    (myKernel.py:159-176), ratio=0 gives nonRotK.K (myKernel.py:255-271).
    """
    dx1, dx2, B11, B12, B22 = _pair_terms(X, X2)
    # the reference takes sqrt then squares (myKernel.py:35,38); kept for rounding parity
    norm = np.sqrt(np.square(dx1) + np.square(dx2))
    rdf2 = l_df * l_df
    Cdf = np.square(norm / l_df)
    aux = 1.0 - Cdf                                    # (p-1) - C with p = 2
    Adf = _blk(B11 / rdf2 + aux, B12 / rdf2, B12 / rdf2, B22 / rdf2 + aux)
    Kdf = np.square(1.0 / l_df) * np.exp(-_blk(Cdf, Cdf, Cdf, Cdf) / 2.0) * Adf
    rcf2 = l_cf * l_cf
    Ccf = np.square(norm / l_cf)
    Acf = _blk(1.0 - B11 / rcf2, -B12 / rcf2, -B12 / rcf2, 1.0 - B22 / rcf2)
    Kcf = np.square(1.0 / l_cf) * np.exp(-_blk(Ccf, Ccf, Ccf, Ccf) / 2.0) * Acf
    return ratio * Kdf + (1.0 - ratio) * Kcf


def helmholtz_Kdiag(M, l_df, l_cf, ratio):
    """Prior variance replicated 2M times (myKernel.py:55-57; M = X.shape[0])."""
    var = ratio * (1.0 / l_df ** 2) + (1.0 - ratio) * (1.0 / l_cf ** 2)
    return np.ones(2 * int(M)) * var


def helmholtz_dK(X, X2, l_df, l_cf, ratio, reference_compat=False):
    """(dK/dl_df, dK/dl_cf, dK/dratio), each [2N,2M].

    reference_compat=True reproduces myKernel.update_gradients_full's integrands
    verbatim (myKernel.py:77-81, 91-96, 99-102), including its wrong length-scale
    terms.  reference_compat=False is the analytically correct derivative
    (SURVEY.md §8(a) row G), checked against central differences in the tests.
    """
    dx1, dx2, B11, B12, B22 = _pair_terms(X, X2)
    r2 = np.square(dx1) + np.square(dx2)
    out = []
    for which, l, w in (("df", l_df, ratio), ("cf", l_cf, 1.0 - ratio)):
        l2, l3, l5 = l * l, l ** 3, l ** 5
        C = r2 / l2
        if which == "df":
            A = _blk(B11 / l2 + 1.0 - C, B12 / l2, B12 / l2, B22 / l2 + 1.0 - C)
            G = _blk(r2 - B11, -B12, -B12, r2 - B22)        # r^2 I - B
        else:
            A = _blk(1.0 - B11 / l2, -B12 / l2, -B12 / l2, 1.0 - B22 / l2)
            G = _blk(B11, B12, B12, B22)                    # B
        C4 = _blk(C, C, C, C)
        E = np.exp(-C4 / 2.0)
        if reference_compat:
            # myKernel.py:81 / :96   dAdf + Adf*(2*l2 - C*l2)/l5
            d = w * E * ((2.0 / l3) * G + A * (2.0 * l2 - C4 * l2) / l5)
        else:
            d = w * E * ((C4 - 2.0) / l3 * A + (2.0 / l5) * G)
        out.append(d)
    Kdf = helmholtz_K(X, X2, l_df, l_cf, 1.0)
    Kcf = helmholtz_K(X, X2, l_df, l_cf, 0.0)
    out.append(Kdf - Kcf)                                   # myKernel.py:99-102
    return tuple(out)


def kernel_grad_sums(dL_dK, X, X2, l_df, l_cf, ratio, reference_compat=False):
    """What update_gradients_full stores in .gradient: sum(dK/dtheta * dL_dK)
    for (length_df, length_cf, ratio)  (myKernel.py:104-106)."""
    d = helmholtz_dK(X, X2, l_df, l_cf, ratio, reference_compat)
    return np.array([np.sum(di * dL_dK) for di in d])


def rbf_ard_K(X, X2, variance, lengthscales):
    """variance * exp(-0.5 * sum_d (dx_d/l_d)^2): the scalar ARD-RBF of
    krig.py:174 (sklearn ``HP[0]*RBF(length_scale=[...])``) / krig.py:388."""
    X = np.asarray(X, dtype=np.float64)
    X2 = X if X2 is None else np.asarray(X2, dtype=np.float64)
    ls = np.asarray(lengthscales, dtype=np.float64)
    a = X / ls
    b = X2 / ls
    d2 = np.zeros((X.shape[0], X2.shape[0]))
    for k in range(X.shape[1]):                 # elementwise differences: no cancellation
        diff = a[:, k][:, None] - b[:, k][None, :]
        d2 += diff * diff
    return variance * np.exp(-0.5 * d2)


# --------------------------------------------------------------------------------------
# fit / predict / log marginal likelihood  (Rasmussen & Williams Alg. 2.1, the algebra of
# sklearn _gpr.py:350-367,446-491,583-656 and of GPy's ExactGaussianInference)
# --------------------------------------------------------------------------------------
def fit(X, y, l_df, l_cf, ratio, noise, jitter=0.0):
    """L = chol(K + (noise+jitter) I), alpha = K^-1 y, LML.

    Reference call sites: GPy GPRegression(X,Y,k) at GP_plots.py:763, krig.py:411;
    numpy K + noise*I at GP_laser.py:113-115,177-179; LML convention of
    sklearn _gpr.py:613-615 (-0.5 y'alpha - sum log L_ii - n/2 log 2pi).
    """
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    K = helmholtz_K(X, None, l_df, l_cf, ratio)
    n = K.shape[0]
    K[np.diag_indices(n)] += noise + jitter
    L = sla.cholesky(K, lower=True, check_finite=False)
    alpha = sla.cho_solve((L, True), y, check_finite=False)
    val = -0.5 * float(y @ alpha) - float(np.sum(np.log(np.diag(L)))) - 0.5 * n * LOG_2PI
    return {"L": L, "alpha": alpha, "lml": val}


def fit_chunked(X, y, l_df, l_cf, ratio, noise, jitter=0.0, chunk=1024):
    """Same result as fit() for large N with bounded memory: the covariance is assembled in row chunks
    of observations (helmholtz_K allocates ~35 temporaries of its output size, myKernel.py:27-53) into ONE
    n x n array that the Cholesky then overwrites.  Peak memory ~ 8 n^2 bytes + chunk temporaries
    (N = 8192: 2.2 GB, N = 16384: 8.6 GB)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    N = X.shape[0]
    n = 2 * N
    K = np.empty((n, n))
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        Kc = helmholtz_K(X[s:e], X, l_df, l_cf, ratio)        # [2c, 2N]: rows (component, observation of the chunk)
        c = e - s
        K[s:e] = Kc[:c]
        K[N + s:N + e] = Kc[c:]
    K[np.diag_indices(n)] += noise + jitter
    L = sla.cholesky(K, lower=True, overwrite_a=True, check_finite=False)
    alpha = sla.cho_solve((L, True), y, check_finite=False)
    val = -0.5 * float(y @ alpha) - float(np.sum(np.log(np.diag(L)))) - 0.5 * n * LOG_2PI
    return {"L": L, "alpha": alpha, "lml": val}


def lml(X, y, l_df, l_cf, ratio, noise, jitter=0.0):
    return fit(X, y, l_df, l_cf, ratio, noise, jitter)["lml"]


def predict(X, fitres, l_df, l_cf, ratio, Xs, noise=0.0, include_noise=False, chunk=4096):
    """Posterior mean [2M] and marginal variance [2M] at Xs.

    mean = K* alpha (sklearn _gpr.py:446-447; GP_scripts.getMean GP_scripts.py:44-46);
    V = L^-1 K*^T, var = diag(K**) - colsumsq(V), negatives clamped to 0
    (_gpr.py:460-462,480-491).  include_noise adds the Gaussian noise variance as
    GPy's model.predict does (krig.py:543-544).  Outputs are component-major:
    [:M] first component, [M:] second (GP_laser.py:184-185).
    """
    Xs = np.asarray(Xs, dtype=np.float64)
    M = Xs.shape[0]
    mean = np.empty(2 * M)
    var = np.empty(2 * M)
    kss = helmholtz_Kdiag(1, l_df, l_cf, ratio)[0]
    L, alpha = fitres["L"], fitres["alpha"]
    for s in range(0, M, chunk):
        e = min(M, s + chunk)
        Ks = helmholtz_K(Xs[s:e], X, l_df, l_cf, ratio)           # [2m, 2N] rows = grid
        mu = Ks @ alpha
        V = sla.solve_triangular(L, Ks.T, lower=True, check_finite=False)
        v = kss - np.einsum("ij,ij->j", V, V)
        m = e - s
        mean[s:e], mean[M + s:M + e] = mu[:m], mu[m:]
        var[s:e], var[M + s:M + e] = v[:m], v[m:]
    var = np.where(var < 0.0, 0.0, var)
    if include_noise:
        var = var + noise
    return mean, var


def lml_and_grad(X, y, l_df, l_cf, ratio, noise, jitter=0.0, reference_compat=False):
    """LML and d LML / d(l_df, l_cf, ratio, noise).

    dL_dK = 0.5 (alpha alpha' - K^-1) (GPy ExactGaussianInference; sklearn
    _gpr.py:629-654), kernel parameters through kernel_grad_sums
    (myKernel.py:104-106), noise through trace(dL_dK).
    """
    f = fit(X, y, l_df, l_cf, ratio, noise, jitter)
    L, alpha = f["L"], f["alpha"]
    n = L.shape[0]
    Kinv = sla.cho_solve((L, True), np.eye(n), check_finite=False)
    dL_dK = 0.5 * (np.outer(alpha, alpha) - Kinv)
    g = kernel_grad_sums(dL_dK, X, None, l_df, l_cf, ratio, reference_compat)
    return f["lml"], np.concatenate([g, [np.trace(dL_dK)]])


def fit_predict_inverse_form(X, y, l_df, l_cf, ratio, noise, Xs, ks_cf_weight=None,
                             want_var=True):
    """The reference's own numpy formulation (explicit inverse).

    K = rate*K_df + (1-rate)*K_cf + noise*I; Ki = inv(K)  (GP_laser.py:113-118,177-180)
    Ks = K(X*, X)                                         (GP_laser.py:122,181)
    mean = Ks Ki y                                        (GP_scripts.py:44-46)
    Cov  = Kss - Ks Ki Ks^T, only its diagonal kept       (GP_laser.py:128-131)
    ks_cf_weight overrides the curl-free weight in K* only: simLaser uses
    (1-rate)*rate there (GP_laser.py:181, a reference defect).
    """
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    K = helmholtz_K(X, None, l_df, l_cf, ratio)
    K = K + np.identity(K.shape[0]) * noise
    Ki = np.linalg.inv(K)
    if ks_cf_weight is None:
        Ks = helmholtz_K(Xs, X, l_df, l_cf, ratio)
    else:
        Ks = (ratio * helmholtz_K(Xs, X, l_df, l_cf, 1.0)
              + ks_cf_weight * helmholtz_K(Xs, X, l_df, l_cf, 0.0))
    mean = np.reshape(np.dot(Ks, np.dot(Ki, y)), [-1])
    if not want_var:
        return mean, None
    kss = helmholtz_Kdiag(1, l_df, l_cf, ratio)[0]
    var = kss - np.einsum("ij,ij->i", Ks @ Ki, Ks)
    return mean, var


# --------------------------------------------------------------------------------------
# space-time product kernel.  Kt.K (myKernel.py:350-360) is a time RBF
# var*exp(-dt^2/2l^2) tiled over the 2x2 blocks; scratch.coKriging multiplies it with the
# divergence-free kernel of (y, x) (scratch.py:506-508: k = kt * kxy, GPy's Prod = elementwise
# product).  The reference code path is dead upstream (ctor keyword mismatch), so this part of the
# oracle restates the formulas only: "parity unpinned".
# --------------------------------------------------------------------------------------
def kt_K(t, t2, var, lengthscale):
    """Kt.K: [[C, C], [C, C]] with C the 1-D RBF of the times (myKernel.py:350-360)."""
    t = np.asarray(t, dtype=np.float64).reshape(-1, 1)
    t2 = t if t2 is None else np.asarray(t2, dtype=np.float64).reshape(-1, 1)
    C = rbf_ard_K(t, t2, var, [lengthscale])
    return _blk(C, C, C, C)


def st_K(X3, X3b, l_df, l_cf, ratio, tvar, lt):
    """Kt(t) * Helmholtz(a, b) for points [N,3] rows (t, a, b), block layout [2N,2M]."""
    X3 = np.asarray(X3, dtype=np.float64)
    X3b = X3 if X3b is None else np.asarray(X3b, dtype=np.float64)
    return kt_K(X3[:, 0], X3b[:, 0], tvar, lt) * helmholtz_K(X3[:, 1:3], X3b[:, 1:3], l_df, l_cf, ratio)


def st_dK(X3, X3b, l_df, l_cf, ratio, tvar, lt):
    """dK/d(l_df, l_cf, ratio, tvar, lt) by the product rule (analytic spatial derivatives)."""
    X3 = np.asarray(X3, dtype=np.float64)
    X3b = X3 if X3b is None else np.asarray(X3b, dtype=np.float64)
    T = kt_K(X3[:, 0], X3b[:, 0], tvar, lt)
    H = helmholtz_K(X3[:, 1:3], X3b[:, 1:3], l_df, l_cf, ratio)
    dH = helmholtz_dK(X3[:, 1:3], X3b[:, 1:3], l_df, l_cf, ratio)
    dt = X3[:, 0][:, None] - X3b[:, 0][None, :]
    dt2 = _blk(dt * dt, dt * dt, dt * dt, dt * dt)
    return [T * dH[0], T * dH[1], T * dH[2], T * H / tvar, T * H * dt2 / lt ** 3]


def st_fit(X3, y, l_df, l_cf, ratio, tvar, lt, noise, jitter=0.0):
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    K = st_K(X3, None, l_df, l_cf, ratio, tvar, lt)
    n = K.shape[0]
    K[np.diag_indices(n)] += noise + jitter
    L = sla.cholesky(K, lower=True, check_finite=False)
    alpha = sla.cho_solve((L, True), y, check_finite=False)
    val = -0.5 * float(y @ alpha) - float(np.sum(np.log(np.diag(L)))) - 0.5 * n * LOG_2PI
    return {"L": L, "alpha": alpha, "lml": val}


def st_predict(X3, fitres, l_df, l_cf, ratio, tvar, lt, Xs3, var_add=0.0):
    Xs3 = np.asarray(Xs3, dtype=np.float64)
    Ks = st_K(Xs3, X3, l_df, l_cf, ratio, tvar, lt)
    mean = Ks @ fitres["alpha"]
    V = sla.solve_triangular(fitres["L"], Ks.T, lower=True, check_finite=False)
    var = tvar * helmholtz_Kdiag(1, l_df, l_cf, ratio)[0] - np.einsum("ij,ij->j", V, V)
    return mean, np.where(var < 0.0, 0.0, var) + var_add


def st_lml_and_grad(X3, y, l_df, l_cf, ratio, tvar, lt, noise, jitter=0.0):
    f = st_fit(X3, y, l_df, l_cf, ratio, tvar, lt, noise, jitter)
    n = f["L"].shape[0]
    Kinv = sla.cho_solve((f["L"], True), np.eye(n), check_finite=False)
    W = 0.5 * (np.outer(f["alpha"], f["alpha"]) - Kinv)
    g = [np.sum(d * W) for d in st_dK(X3, None, l_df, l_cf, ratio, tvar, lt)]
    return f["lml"], np.array(g + [np.trace(W)])


# --------------------------------------------------------------------------------------
# sum of space-time Helmholtz terms: what krig.kriging(kernelType = 2, 3, 4, nKernels) asks the
# module myKernel2 for (krig.py:396-407): divFreeK / curlFreeK(input_dim=3, var, lt, ly, lx), their
# sum, nKernels copies with independent parameters.  myKernel2 is NOT in the reference repository,
# so the anisotropic / time-dependent formula is specified here ("parity unpinned"): the stream
# function / potential construction of myKernel.py:39-52 applied to
#   s = var exp(-dt^2/2lt^2 - da^2/2la^2 - db^2/2lb^2).
# What IS pinned: with la == lb, no time and {div-free var=ratio} + {curl-free var=1-ratio} it must
# equal helmholtz_K (the golden-checked restatement of myKernel.py:27-53), and with a shared lt it
# must equal st_K; tests/test_oracle_golden.py checks both.
# --------------------------------------------------------------------------------------
def _hsum_split(X):
    X = np.asarray(X, dtype=np.float64)
    if X.shape[1] == 3:
        return X[:, 0], X[:, 1], X[:, 2]
    return np.zeros(X.shape[0]), X[:, 0], X[:, 1]


def _hsum_term(dt, d1, d2, ty, var, lt, la, lb, has_t):
    """(k, P, R, C): scalar envelope and the (1 - s1)/la^2, (1 - s2)/lb^2, +-d1 d2/(la lb)^2 factors."""
    a1, a2 = 1.0 / la ** 2, 1.0 / lb ** 2
    s1, s2 = d1 * d1 * a1, d2 * d2 * a2
    e = 0.5 * (s1 + s2)
    if has_t:
        e = e + 0.5 * dt * dt / lt ** 2
    k = var * np.exp(-e)
    sgn = -1.0 if ty else 1.0
    return k, a1 * (1.0 - s1), a2 * (1.0 - s2), sgn * d1 * d2 * a1 * a2, s1, s2


def hsum_K(X, X2, types, params):
    """[2N,2M] block matrix; types[q] in {0 div-free, 1 curl-free}; params[q] = (var, lt, la, lb)."""
    t, a, b = _hsum_split(X)
    t2, a2_, b2_ = (t, a, b) if X2 is None else _hsum_split(X2)
    has_t = np.asarray(X).shape[1] == 3
    dt, d1, d2 = t[:, None] - t2[None, :], a[:, None] - a2_[None, :], b[:, None] - b2_[None, :]
    K = 0.0
    for ty, (var, lt, la, lb) in zip(types, np.atleast_2d(params)):
        k, P, R, C, _, _ = _hsum_term(dt, d1, d2, ty, var, lt, la, lb, has_t)
        K = K + (_blk(k * P, k * C, k * C, k * R) if ty else _blk(k * R, k * C, k * C, k * P))
    return K


def hsum_Kdiag(M, types, params):
    """Prior variances, first M entries component 0."""
    v0 = v1 = 0.0
    for ty, (var, lt, la, lb) in zip(types, np.atleast_2d(params)):
        v0 += var / (la ** 2 if ty else lb ** 2)
        v1 += var / (lb ** 2 if ty else la ** 2)
    return np.concatenate([np.full(M, v0), np.full(M, v1)])


def hsum_dK(X, X2, types, params):
    """List over terms of [dK/dvar, dK/dlt, dK/dla, dK/dlb] (analytic)."""
    t, a, b = _hsum_split(X)
    t2, a2_, b2_ = (t, a, b) if X2 is None else _hsum_split(X2)
    has_t = np.asarray(X).shape[1] == 3
    dt, d1, d2 = t[:, None] - t2[None, :], a[:, None] - a2_[None, :], b[:, None] - b2_[None, :]
    out = []
    for ty, (var, lt, la, lb) in zip(types, np.atleast_2d(params)):
        k, P, R, C, s1, s2 = _hsum_term(dt, d1, d2, ty, var, lt, la, lb, has_t)

        def blk(p, r, c):
            return _blk(p, c, c, r) if ty else _blk(r, c, c, p)
        E = k / var if var != 0.0 else _hsum_term(dt, d1, d2, ty, 1.0, lt, la, lb, has_t)[0]
        dvar = blk(E * P, E * R, E * C)
        tt = (dt * dt / lt ** 3) if has_t else np.zeros_like(dt)
        dlt = blk(k * P * tt, k * R * tt, k * C * tt)
        # d(kP)/dla = k/la [s1 P + (4 s1 - 2)/la^2]; d(kR)/dla = k/la s1 R; d(kC)/dla = k/la (s1 - 2) C
        dla = blk(k / la * (s1 * P + (4.0 * s1 - 2.0) / la ** 2), k / la * s1 * R, k / la * (s1 - 2.0) * C)
        dlb = blk(k / lb * s2 * P, k / lb * (s2 * R + (4.0 * s2 - 2.0) / lb ** 2), k / lb * (s2 - 2.0) * C)
        out.append([dvar, dlt, dla, dlb])
    return out


def hsum_kernel_grad_sums(dL_dK, X, X2, types, params):
    return np.array([[np.sum(d * dL_dK) for d in term] for term in hsum_dK(X, X2, types, params)])


def hsum_fit(X, y, types, params, noise, jitter=0.0):
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    K = hsum_K(X, None, types, params)
    n = K.shape[0]
    K[np.diag_indices(n)] += noise + jitter
    L = sla.cholesky(K, lower=True, check_finite=False)
    alpha = sla.cho_solve((L, True), y, check_finite=False)
    val = -0.5 * float(y @ alpha) - float(np.sum(np.log(np.diag(L)))) - 0.5 * n * LOG_2PI
    return {"L": L, "alpha": alpha, "lml": val}


def hsum_predict(X, fitres, types, params, Xs, var_add=0.0):
    Xs = np.asarray(Xs, dtype=np.float64)
    Ks = hsum_K(Xs, X, types, params)
    mean = Ks @ fitres["alpha"]
    V = sla.solve_triangular(fitres["L"], Ks.T, lower=True, check_finite=False)
    var = hsum_Kdiag(Xs.shape[0], types, params) - np.einsum("ij,ij->j", V, V)
    return mean, np.where(var < 0.0, 0.0, var) + var_add


def hsum_lml_and_grad(X, y, types, params, noise, jitter=0.0):
    """(LML, [d/d(var, lt, la, lb)_q ..., d/dnoise])."""
    f = hsum_fit(X, y, types, params, noise, jitter)
    n = f["L"].shape[0]
    Kinv = sla.cho_solve((f["L"], True), np.eye(n), check_finite=False)
    W = 0.5 * (np.outer(f["alpha"], f["alpha"]) - Kinv)
    g = hsum_kernel_grad_sums(W, X, None, types, params).reshape(-1)
    return f["lml"], np.concatenate([g, [np.trace(W)]])


# --------------------------------------------------------------------------------------
# scalar ARD-RBF sum family: the reference's production kernels.  GPy.kern.RBF(input_dim=3,
# ARD=True) summed nKernels times (krig.py:388,405-407); sklearn HP[0]*RBF([..]) + HP[4]*RBF([..])
# + WhiteKernel(noise) (krig.py:174-181).  Pinned against live scikit-learn in
# tests/golden/make_golden_sklearn.py / tests/test_oracle_golden.py.
# --------------------------------------------------------------------------------------
def rbf_sum_K(X, X2, variances, lengthscales):
    """sum_q variances[q] * exp(-1/2 sum_d ((x_d - x'_d)/lengthscales[q][d])^2)."""
    ls = np.atleast_2d(np.asarray(lengthscales, dtype=np.float64))
    K = 0.0
    for v, l in zip(np.atleast_1d(variances), ls):
        K = K + rbf_ard_K(X, X2, float(v), l)
    return K


def rbf_sum_dK(X, X2, variances, lengthscales):
    """List of dK/dtheta, theta ordered (variance_q, lengthscale_q[0..D-1]) per component: GPy's
    RBF ARD gradients dK/dvar = K_q/var, dK/dl_d = K_q * (x_d - x'_d)^2 / l_d^3."""
    X = np.asarray(X, dtype=np.float64)
    X2 = X if X2 is None else np.asarray(X2, dtype=np.float64)
    ls = np.atleast_2d(np.asarray(lengthscales, dtype=np.float64))
    out = []
    for v, l in zip(np.atleast_1d(variances), ls):
        Kq = rbf_ard_K(X, X2, float(v), l)
        out.append(Kq / float(v))
        for d in range(X.shape[1]):
            diff = X[:, d][:, None] - X2[:, d][None, :]
            out.append(Kq * diff * diff / l[d] ** 3)
    return out


def rbf_kernel_grad_sums(dL_dK, X, X2, variances, lengthscales):
    return np.array([np.sum(d * dL_dK) for d in rbf_sum_dK(X, X2, variances, lengthscales)])


def rbf_fit(X, y, variances, lengthscales, noise, jitter=0.0):
    """Cholesky fit of the scalar GP (sklearn _gpr.py:350-367 with alpha=jitter; GPy adds 1e-8)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    K = rbf_sum_K(X, None, variances, lengthscales)
    n = K.shape[0]
    K[np.diag_indices(n)] += noise + jitter
    L = sla.cholesky(K, lower=True, check_finite=False)
    alpha = sla.cho_solve((L, True), y, check_finite=False)
    val = -0.5 * float(y @ alpha) - float(np.sum(np.log(np.diag(L)))) - 0.5 * n * LOG_2PI
    return {"L": L, "alpha": alpha, "lml": val}


def rbf_predict(X, fitres, variances, lengthscales, Xs, var_add=0.0, chunk=4096):
    """mean = K* alpha, var = k** - colsumsq(L^-1 K*^T) clamped at 0, + var_add (the WhiteKernel /
    Gaussian-noise variance both sklearn's predict(return_std) and GPy's predict include)."""
    Xs = np.asarray(Xs, dtype=np.float64)
    M = Xs.shape[0]
    mean, var = np.empty(M), np.empty(M)
    kss = float(np.sum(variances))
    for s in range(0, M, chunk):
        e = min(M, s + chunk)
        Ks = rbf_sum_K(Xs[s:e], X, variances, lengthscales)
        mean[s:e] = Ks @ fitres["alpha"]
        V = sla.solve_triangular(fitres["L"], Ks.T, lower=True, check_finite=False)
        var[s:e] = kss - np.einsum("ij,ij->j", V, V)
    return mean, np.where(var < 0.0, 0.0, var) + var_add


def rbf_lml_and_grad(X, y, variances, lengthscales, noise, jitter=0.0):
    """LML and its gradient over (variance_q, lengthscale_q[..])_q then noise."""
    f = rbf_fit(X, y, variances, lengthscales, noise, jitter)
    n = f["L"].shape[0]
    Kinv = sla.cho_solve((f["L"], True), np.eye(n), check_finite=False)
    dL_dK = 0.5 * (np.outer(f["alpha"], f["alpha"]) - Kinv)
    g = rbf_kernel_grad_sums(dL_dK, X, None, variances, lengthscales)
    return f["lml"], np.concatenate([g, [np.trace(dL_dK)]])


# --------------------------------------------------------------------------------------
# config-1 data preparation (host side; GP_laser.py:147-175)
# --------------------------------------------------------------------------------------
def haversine_km(lat1, lon1, lat2, lon2, radius_km=6371.009):
    """Great-circle distance on a sphere of geopy's default radius; stands in for
    geopy.distance.GreatCircleDistance(...).km (GP_laser.py:166-167), geopy absent."""
    p1, p2 = np.radians(lat1), np.radians(lat2)
    dphi = p2 - p1
    dlmb = np.radians(np.asarray(lon2) - np.asarray(lon1))
    a = np.sin(dphi / 2.0) ** 2 + np.cos(p1) * np.cos(p2) * np.sin(dlmb / 2.0) ** 2
    return 2.0 * radius_km * np.arcsin(np.sqrt(a))


def simlaser_inputs(tr, ts=0, lat0=28.69, lon0=-88.28, dx=0.5):
    """Observation coordinates/velocities and the 51x51 grid of GP_laser.simLaser
    (GP_laser.py:147-175).  Distances are unsigned, like the reference's."""
    lat = np.asarray(tr.lat)[:, ts]
    lon = np.asarray(tr.lon)[:, ts]
    xo = haversine_km(lat, lon, lat, np.full_like(lon, lon0))
    yo = haversine_km(lat, lon, np.full_like(lat, lat0), lon)
    uo = np.asarray(tr.u)[:, ts]
    vo = np.asarray(tr.v)[:, ts]
    x = np.arange(0, 25 + dx, dx)
    yv = np.arange(0, 25 + dx, dx)
    Xg, Yg = np.meshgrid(x, yv)
    Xs = np.stack([Xg.reshape(-1), Yg.reshape(-1)], axis=1)
    return np.stack([xo, yo], axis=1), np.concatenate([uo, vo]), Xs


# --------------------------------------------------------------------------------------
# spatial order of the observations inside a fit (no counterpart upstream: the reference takes the
# drifters in file order, GP_laser.py:62-110; a GP does not depend on the order)
# --------------------------------------------------------------------------------------
def morton_order(P):
    """Permutation along a Z-order curve: 8 bits per axis on a square raster of the bounding box (cell =
    extent / 255.999 of the longer side), x bits even, y bits odd, ties in the given order.  What
    2d-gp_b200/csrc/order.cu computes on the device; internal observation i is the caller's perm[i]."""
    P = np.asarray(P, dtype=np.float64)
    lo = P.min(axis=0)
    ext = max(P[:, 0].max() - lo[0], P[:, 1].max() - lo[1])
    s = 255.999 / ext if (ext > 0 and np.isfinite(ext)) else 0.0
    q = np.minimum(np.maximum((P - lo) * s, 0.0), 255.0).astype(np.uint32)

    def spread(v):
        v = (v | (v << 4)) & 0x0F0F
        v = (v | (v << 2)) & 0x3333
        v = (v | (v << 1)) & 0x5555
        return v
    key = spread(q[:, 0]) | (spread(q[:, 1]) << 1)
    return np.argsort(key, kind="stable")

