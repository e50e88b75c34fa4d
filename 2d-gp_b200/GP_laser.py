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


class _TracksUnpickler(pickle.Unpickler):
    """Unpickler for the reference's track pickles that imports nothing from the data's directory: the
    container classes of laser_class.py map onto this package's own (laser_io_methods), numpy arrays /
    dtypes / scalars and datetimes are rebuilt, every other global is refused."""
    _NUMPY = {"_reconstruct", "ndarray", "dtype", "scalar", "_frombuffer"}

    def find_class(self, module, name):
        if module in ("laser_class", "laser_io_methods") and name in ("interpolated_tracks", "drifter"):
            from . import laser_io_methods
            return getattr(laser_io_methods, name)
        if module.split(".")[0] == "numpy" and name in self._NUMPY:
            import importlib
            try:
                return getattr(importlib.import_module(module), name)
            except (ImportError, AttributeError):          # numpy.core -> numpy._core renames
                import numpy._core.multiarray as ma
                return getattr(ma, name) if hasattr(ma, name) else getattr(np, name)
        if (module, name) in (("datetime", "datetime"), ("datetime", "timedelta"), ("datetime", "date")):
            import datetime
            return getattr(datetime, name)
        if (module, name) in (("copy_reg", "_reconstructor"), ("copyreg", "_reconstructor")):
            import copyreg
            return copyreg._reconstructor
        if (module, name) in (("__builtin__", "object"), ("builtins", "object")):
            return object
        raise pickle.UnpicklingError("refusing to load %s.%s from a track pickle" % (module, name))


def load_tracks(path="simulTracks.pkl", class_dir=None):
    """Read an ``interpolated_tracks`` object written by the reference (Python 2 pickle) with a
    restricted unpickler: no module is imported from the pickle's directory (``class_dir`` is accepted
    for compatibility and ignored), only the track containers, numpy arrays and datetimes are rebuilt."""
    with open(path, "rb") as f:
        return _TracksUnpickler(f, encoding="latin1").load()


def laser(ts=20, nsteps=8, l_df=5, l_cf=5, tau=1, rate=0.5, noise=0.0025, nsamples=1, tracks=None,
          path="interp_ALL_2016_2_7.pkl"):
    """GP_laser.laser (GP_laser.py:16-142) on the GPU: LASER drifter tracks (positions, velocities,
    drogue status on a 15-minute grid) -> observations of ``nsteps`` time steps from ``ts`` ->
    every third one kept for the fit, the rest for verification -> posterior mean and variance on
    a 0.5 km grid and the mean at the verification points.

    Host side as in the reference: undrogued samples and |u|,|v| > 2 m/s dropped, EPSG:3452
    projection to km, origin shifted to (2, 2) km.  Device side: one fit, one fused predictive pass
    for the grid (mean and marginal variance: the reference forms the full 2M x 2M ``Kss`` and
    ``Cov`` only to take their diagonals, GP_laser.py:128-131) and one for the verification points.
    ``tracks`` may carry the unpickled ``interpolated_tracks`` object (the LASER pickle is not
    shipped upstream).  Returns x, y, uf, vf, xo, yo, uo, vo, uvar, vvar, xt, yt, ut, vt, uft, vft."""
    from .projection import NAD83
    tr = tracks if tracks is not None else load_tracks(path, os.path.dirname(os.path.abspath(path)))
    st = 0
    et = st + (96 * 7) + 1
    latt, lont = np.array(tr.lat)[:, st:et], np.array(tr.lon)[:, st:et]
    drogue = np.array(tr.drogueStat)[:, st:et]
    uob, vob = np.array(tr.u, dtype=np.float64)[:, st:et], np.array(tr.v, dtype=np.float64)[:, st:et]
    uob[np.where(drogue == 0)] = np.nan
    vob[np.where(drogue == 0)] = np.nan
    xob, yob = NAD83(lont, latt)                       # NaN positions stay NaN
    xo = np.reshape(xob[:, ts:ts + nsteps], [-1]) / 1000.
    yo = np.reshape(yob[:, ts:ts + nsteps], [-1]) / 1000.
    uo = np.reshape(uob[:, ts:ts + nsteps], [-1])
    vo = np.reshape(vob[:, ts:ts + nsteps], [-1])
    with np.errstate(invalid="ignore"):
        uo[np.where(np.abs(uo) > 2)] = np.nan
        vo[np.where(np.abs(vo) > 2)] = np.nan
    ok = np.where((~np.isnan(uo)) & (~np.isnan(vo)) & (~np.isnan(xo)) & (~np.isnan(yo)))
    uo, vo, xo, yo = uo[ok], vo[ok], xo[ok], yo[ok]
    xo = xo - xo.min() + 2
    yo = yo - yo.min() + 2
    if nsamples > 0:
        samples = np.arange(0, xo.size, 3)
        test = np.array(sorted(set(range(xo.size)) - set(samples)), dtype=int)
        xt, yt, ut, vt = xo[test], yo[test], uo[test], vo[test]
        xo, yo, uo, vo = xo[samples], yo[samples], uo[samples], vo[samples]
    else:
        xt = yt = ut = vt = np.array([0])
    dx = 0.5
    x = np.arange(np.min([xo.min(), xt.min()]) - 5, np.max([xo.max(), xt.max()]) + 5, dx)
    y = np.arange(np.min([yo.min(), yt.min()]) - 5, np.max([yo.max(), yt.max()]) + 5, dx)
    X, Y = np.meshgrid(x, y)
    Xs = np.stack([np.reshape(X, [X.size]), np.reshape(Y, [Y.size])], axis=1)
    gp = HelmholtzGP(np.stack([xo, yo], axis=1), np.concatenate([uo, vo]), l_df, l_cf, rate, noise)
    gp.fit()
    mean, var = gp.predict(Xs)
    f, var = mean.cpu().numpy(), var.cpu().numpy()
    uf, vf = np.reshape(f[:f.size // 2], [y.size, -1]), np.reshape(f[f.size // 2:], [y.size, -1])
    uvar, vvar = np.reshape(var[:X.size], [y.size, -1]), np.reshape(var[X.size:], [y.size, -1])
    ft = gp.predict(np.stack([xt, yt], axis=1))[0].cpu().numpy()
    return x, y, uf, vf, xo, yo, uo, vo, uvar, vvar, xt, yt, ut, vt, ft[:ft.size // 2], ft[ft.size // 2:]


laser2 = laser          # GP_laser.laser2 (GP_laser.py:195-321) is a verbatim copy of laser upstream


def simLaser(ts=0, l_df=2, l_cf=2, rate=0.5, noise=0.05, tracks=None, path="simulTracks.pkl",
             simlaser_compat=False, return_var=False):
    """Returns X, Y, uf, vf, xob, yob, u, v with the reference's shapes and ordering (plus uvar, vvar on
    request).

    DEVIATION FROM UPSTREAM BY DEFAULT: the reference's simLaser weights the cross-covariance as
    rate*K*_df + (1-rate)*rate*K*_cf (GP_laser.py:181, an extra ``rate`` on the curl-free part that its
    own K, GP_laser.py:177-179, and laser(), GP_laser.py:122, do not have), so its (uf, vf) are not the
    posterior mean of the model it fits.  The default here is the consistent weighting of GP_laser.py:122.
    Pass ``simlaser_compat=True`` for a drop-in comparison with upstream output: it reproduces
    GP_laser.py:181 exactly (golden-tested both ways in tests/test_reference_surface.py and
    tests/test_gpu_parity.py::test_simlaser_fit_predict); the variance is then not returned, as the
    reference's own variance uses yet another weighting."""
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
