"""Host-side map projection used by krig (pyproj is not required).

The reference projects lon/lat with ``pyproj.Proj("+init=EPSG:3452")`` (krig.py:19-20,291):
NAD83 / Louisiana South, a Lambert conformal conic (2SP) on GRS80 with standard parallels
29.3 N and 30.7 N, origin 28.5 N / 91.3333 W, false easting 1 000 000 m.  Old pyproj returns
metres for ``+init=`` strings (preserve_units=False), which is what krig divides by 1000 to get
km.  Snyder, "Map Projections -- A Working Manual", eqs. 15-1 .. 15-9.
"""
from __future__ import annotations

import numpy as np

_A = 6378137.0
_F = 1.0 / 298.257222101
_E = np.sqrt(2 * _F - _F * _F)
_LAT1, _LAT2, _LAT0, _LON0 = np.radians([30.7, 29.3, 28.5, -(91.0 + 20.0 / 60.0)])
_X0, _Y0 = 1000000.0, 0.0


def _m(phi):
    return np.cos(phi) / np.sqrt(1 - (_E * np.sin(phi)) ** 2)


def _t(phi):
    s = _E * np.sin(phi)
    return np.tan(np.pi / 4 - phi / 2) / ((1 - s) / (1 + s)) ** (_E / 2)


_N = (np.log(_m(_LAT1)) - np.log(_m(_LAT2))) / (np.log(_t(_LAT1)) - np.log(_t(_LAT2)))
_FF = _m(_LAT1) / (_N * _t(_LAT1) ** _N)
_RHO0 = _A * _FF * _t(_LAT0) ** _N


def NAD83(lon, lat):
    """(lon, lat) in degrees -> (x, y) in metres, EPSG:3452 geometry.  NaN in, NaN out."""
    lon = np.asarray(lon, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    rho = _A * _FF * _t(np.radians(lat)) ** _N
    theta = _N * (np.radians(lon) - _LON0)
    return _X0 + rho * np.sin(theta), _Y0 + _RHO0 - rho * np.cos(theta)
