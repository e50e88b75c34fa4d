"""GP_laser.simLaser (GP_laser.py:145-187) on the GPU: config 1 of BASELINE.json.

Host side (as in the reference): unpickle simulTracks.pkl, great-circle distances to
(28.69, -88.28) in km, pick snapshot ``ts``, 51 x 51 grid at 0.5 km.  Device side: fit
(covariance build, Cholesky / L^-1, alpha) and the fused predictive pass.
"""
from __future__ import annotations

import os
import pickle
import sys

import numpy as np

from .engine import HelmholtzGP

EARTH_RADIUS_KM = 6371.009       # geopy's default sphere (geopy.distance.GreatCircleDistance)


def great_circle_km(lat1, lon1, lat2, lon2):
    """Haversine stand-in for geopy's GreatCircleDistance(...).km  (GP_laser.py:166-167)."""
    p1, p2 = np.radians(lat1), np.radians(lat2)
    a = np.sin((p2 - p1) / 2.0) ** 2 + np.cos(p1) * np.cos(p2) * np.sin(np.radians(lon2 - lon1) / 2.0) ** 2
    return 2.0 * EARTH_RADIUS_KM * np.arcsin(np.sqrt(a))


def load_tracks(path="simulTracks.pkl", class_dir=None):
    """Unpickle an ``interpolated_tracks`` object written by the reference (Python 2 pickle).
    ``class_dir`` is the directory holding the reference's laser_class.py."""
    if class_dir and class_dir not in sys.path:
        sys.path.insert(0, class_dir)
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")


def simLaser(ts=0, l_df=2, l_cf=2, rate=0.5, noise=0.05, tracks=None, path="simulTracks.pkl",
             simlaser_compat=False, return_var=False):
    """Returns X, Y, uf, vf, xob, yob, u, v like the reference (plus uvar, vvar on request).

    simlaser_compat=True reproduces the reference's K* weighting (1-rate)*rate on the
    curl-free part (GP_laser.py:181); the default uses the weighting of GP_laser.py:122.
    """
    dx = 0.5
    x = np.arange(0, 25 + dx, dx)
    y = np.arange(0, 25 + dx, dx)
    X, Y = np.meshgrid(x, y)
    Xs = np.stack([np.reshape(X, [X.size]), np.reshape(Y, [Y.size])], axis=1)
    tr = tracks if tracks is not None else load_tracks(path, os.path.dirname(os.path.abspath(path)))
    lat0, lon0 = 28.69, -88.28
    lat, lon = np.asarray(tr.lat), np.asarray(tr.lon)
    xob = great_circle_km(lat, lon, lat, np.full_like(lon, lon0))
    yob = great_circle_km(lat, lon, np.full_like(lat, lat0), lon)
    xo, yo = xob[:, ts], yob[:, ts]
    obs = np.concatenate([np.asarray(tr.u)[:, ts], np.asarray(tr.v)[:, ts]])
    gp = HelmholtzGP(np.stack([xo, yo], axis=1), obs, l_df, l_cf, rate, noise)
    gp.fit()
    if simlaser_compat:
        # K* = rate*K*_df + (1-rate)*rate*K*_cf: two passes with single-component kernels
        # against the same alpha (gp2d_predict takes theta separately from the fit state)
        gp.ratio = 1.0
        m_df, _ = gp.predict(Xs)
        gp.ratio = 0.0
        m_cf, _ = gp.predict(Xs)
        gp.ratio = float(rate)
        f = (rate * m_df + (1 - rate) * rate * m_cf).cpu().numpy()
        var = None
    else:
        mean, var = gp.predict(Xs)
        f, var = mean.cpu().numpy(), var.cpu().numpy()
    uf = np.reshape(f[:f.size // 2], [y.size, -1])
    vf = np.reshape(f[f.size // 2:], [y.size, -1])
    if return_var and var is not None:
        return (X, Y, uf, vf, xob, yob, tr.u, tr.v,
                np.reshape(var[:var.size // 2], [y.size, -1]), np.reshape(var[var.size // 2:], [y.size, -1]))
    return X, Y, uf, vf, xob, yob, tr.u, tr.v
