"""createNC / writeNC with the reference's file layout (printNCFiles.py:5-44) on
scipy.io.netcdf_file (NetCDF-3 64-bit offset), so netCDF4 is not required.  Host I/O."""
from __future__ import annotations

import numpy as np
from scipy.io import netcdf_file


def createNC(outFile, T, Y, X, hyp):
    NY, NX, NH = np.size(Y), np.size(X), np.size(hyp)
    f = netcdf_file(outFile, "w", version=2)
    f.createDimension("time", None)
    f.createDimension("y", NY)
    f.createDimension("x", NX)
    f.createDimension("hyperparam", NH)
    y = f.createVariable("y", "f4", ("y",))
    x = f.createVariable("x", "f4", ("x",))
    times = f.createVariable("time", "f4", ("time",))
    f.createVariable("hyperparam_u", "f4", ("hyperparam",))
    f.createVariable("hyperparam_v", "f4", ("hyperparam",))
    rec = [f.createVariable(name, "f4", ("time", "y", "x")) for name in ("v", "u", "vvar", "uvar")]
    y[:] = np.squeeze(Y)
    x[:] = np.squeeze(X)
    NT = np.size(T)
    times[:NT] = np.atleast_1d(np.squeeze(T))
    for var in rec:                      # every record variable must hold the same number of records
        var[:NT] = np.zeros((NT, NY, NX), dtype=np.float32)
    f.close()


def writeNC(f, varname, data):
    NT = np.size(data, 0)
    var = f.variables[varname]
    if len(var.shape) == 1:
        var[0:NT] = data
    elif len(var.shape) == 3:
        var[0:NT, :, :] = data
    return f


def openNC(path, mode="a"):
    return netcdf_file(path, mode, version=2, mmap=False)
