"""Function-form call surface of the reference's GP_scripts.py, computed on the GPU.

Same names, argument order and output layouts (component-major blocks, GP_scripts.py:89-95);
numpy in, numpy out.  The O(N^2) Python loops of compute_K / compute_Ks (GP_scripts.py:74-123)
become one covariance-build kernel launch; the explicit inverse and full posterior covariance
of getCov are kept for drop-in use at small sizes (the production path is
``HelmholtzGP.predict``, which never forms them).
"""
from __future__ import annotations

import numpy as np

from . import engine


def _pts(x1, x2):
    return np.stack([np.reshape(np.asarray(x1, dtype=np.float64), [-1]),
                     np.reshape(np.asarray(x2, dtype=np.float64), [-1])], axis=1)


def myKernel(xa, xb, r_df, r_cf, alpha=1):
    """alpha*K_divfree + (1-alpha)*K_curlfree, [2N,2M]  (GP_scripts.py:6-42)."""
    return engine.kernel_K(np.asarray(xa, dtype=np.float64), np.asarray(xb, dtype=np.float64),
                           float(r_df), float(r_cf), float(alpha)).cpu().numpy()


def _ratio(divFree):
    if divFree == 1:
        return 1.0
    if divFree == 2:
        return 0.0
    raise ValueError("divFree must be 1 (divergence-free) or 2 (curl-free)")


def nonDivK(xa, xb, sigma, divFree=1):
    """2x2 covariance block between two points; any other divFree gives the scalar
    squared-exponential exp(-|xa-xb|^2 / 2 sigma^2) (GP_scripts.py:57-69)."""
    xa = np.asarray(xa, dtype=np.float64).reshape(1, 2)
    xb = np.asarray(xb, dtype=np.float64).reshape(1, 2)
    if divFree not in (1, 2):
        return float(engine.rbf_K(xa, xb, [1.0], [[float(sigma)] * 2]).cpu().numpy()[0, 0])
    return engine.kernel_K(xa, xb, float(sigma), float(sigma), _ratio(divFree)).cpu().numpy()


def sqExp(x1, y1, x2, y2, sigma):
    """Scalar squared-exponential between the points (x1, y1) and (x2, y2), [I,J]  (GP_scripts.py:125-134)."""
    return engine.rbf_K(_pts(x1, y1), _pts(x2, y2), [1.0], [[float(sigma)] * 2]).cpu().numpy()


def rbf(x1, x2, l=1, sigma=1, noise=0):
    """sigma^2 exp(-(x2-x1)^2 / 2 l^2), plus noise on the diagonal when the sizes match
    (GP_scripts.py:136-142)."""
    a = np.reshape(np.asarray(x1, dtype=np.float64), [-1, 1])
    b = np.reshape(np.asarray(x2, dtype=np.float64), [-1, 1])
    K = engine.rbf_K(a, b, [float(sigma) ** 2], [[float(l)]]).cpu().numpy()
    if a.size == b.size:
        K = K + np.identity(a.size) * noise
    return K


def compute_K(x1, x2, sigma, divFree=1):
    """K at the sample inputs, [2N,2N]  (GP_scripts.py:74-95)."""
    return engine.kernel_K(_pts(x1, x2), None, float(sigma), float(sigma), _ratio(divFree)).cpu().numpy()


def compute_Ks(x1, x2, x1s, x2s, sigma, divFree=1):
    """K(X*, X), [2M,2N]: rows are grid points  (GP_scripts.py:97-123)."""
    return engine.kernel_K(_pts(x1s, x2s), _pts(x1, x2), float(sigma), float(sigma), _ratio(divFree)).cpu().numpy()


def getMean(KS, Ki, y):
    """f = KS (Ki y)  (GP_scripts.py:44-46), two DMMA GEMMs on the device."""
    a = engine.matmul(Ki, np.reshape(np.asarray(y, dtype=np.float64), [-1]))
    return engine.matmul(KS, a).cpu().numpy().reshape(-1)


def getCov(x1, x2, x1s, x2s, sigma=0.2, divFree=1):
    """(Kss - Ks Ki Ks^T, Ki, Ks)  (GP_scripts.py:48-54).  Full [2M,2M] posterior covariance:
    O(M^2) memory, kept for drop-in use at the reference's own (small) sizes."""
    X, Xs = _pts(x1, x2), _pts(x1s, x2s)
    r = _ratio(divFree)
    Ki = engine.spd_inverse(engine.kernel_K(X, None, float(sigma), float(sigma), r))
    Ks = engine.kernel_K(Xs, X, float(sigma), float(sigma), r)
    Kss = engine.kernel_K(Xs, None, float(sigma), float(sigma), r)
    ML = Kss - engine.matmul(engine.matmul(Ks, Ki), Ks.t().contiguous())
    return ML.cpu().numpy(), Ki.cpu().numpy(), Ks.cpu().numpy()


# ---- host-side helpers kept from the reference (metrics / synthetic test field) ------------
def generate_2D_gaussian(divFree=1):
    """Synthetic divergence-free (or curl-free) velocity field  (GP_scripts.py:202-223)."""
    dx = 0.05
    x = np.arange(-1, 1 + dx, dx)
    y = np.arange(-1, 1 + dx, dx)
    A, l = 1.5, 3.
    X, Y = np.meshgrid(x, y)
    phi = A * np.exp(-(X ** 2) / l - (Y ** 2) / l)
    dphi_dy = np.diff(phi, axis=0)
    dphi_dx = np.diff(phi, axis=1)
    if divFree == 1:
        um = (dphi_dy[:, :-1] + dphi_dy[:, 1:]) / 2.
        vm = -(dphi_dx[:-1, :] + dphi_dx[1:, :]) / 2.
    else:
        um = -(dphi_dx[:-1, :] + dphi_dx[1:, :]) / 2.
        vm = -(dphi_dy[:, :-1] + dphi_dy[:, 1:]) / 2.
    xm = (x[:-1] + x[1:]) / 2.
    ym = (x[:-1] + x[1:]) / 2.
    return x, y, phi, xm, ym, um, vm


def rmse1(ys, y):
    """Root-mean-square error (GP_scripts.py:172-176)."""
    error = np.reshape(np.asarray(ys) - np.asarray(y), [-1])
    return np.sqrt(np.mean(np.square(error)))


def rmse(x1s, x2s, f1, f2, x1, x2, y1, y2, knd=''):
    """RMSE of both components at the data points (GP_scripts.py:144-170).  Only the direct
    comparison (knd='') and radial-basis interpolation (knd='rbf') are kept: scipy removed
    interp2d, which the 'cubic' / 'linear' branches used."""
    if knd == 'rbf':
        from scipy.interpolate import Rbf
        y1s, y2s = Rbf(x1s, x2s, f1)(x1, x2), Rbf(x1s, x2s, f2)(x1, x2)
    elif knd in ('cubic', 'linear'):
        raise NotImplementedError("scipy.interpolate.interp2d no longer exists")
    else:
        y1s, y2s = f1, f2
    r1, r2 = rmse1(y1s, y1), rmse1(y2s, y2)
    print(r1 / (np.max(y1) - np.min(y1)), r2 / (np.max(y2) - np.min(y2)))
    return r1, r2


def absoluteError(y, f, x1f, x2f, x1, x2):
    """|y - f| and the distance of each grid point to the nearest observation (GP_scripts.py:178-200)."""
    f, y = np.reshape(f, [-1]), np.reshape(y, [-1])
    X1f, X1 = np.meshgrid(x1f, x1)
    X2f, X2 = np.meshgrid(x2f, x2)
    return np.abs(y - f), np.min(np.sqrt(np.square(X1 - X1f) + np.square(X2 - X2f)), 0)


def vel_grad(x, y, u, v):
    """Finite-difference divergence  (GP_scripts.py:226-233)."""
    dx, dy = np.diff(x), np.diff(y)
    du, dv = np.diff(u, axis=1), np.diff(v, axis=0)
    dum = (du[1:, :] + du[:-1, :]) / 2.
    dvm = (dv[:, 1:] + dv[:, :-1]) / 2.
    return dum / dx + dvm / dy
