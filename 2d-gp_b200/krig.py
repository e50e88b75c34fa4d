"""Kriging workflows with the reference's call surface (krig.py), on the CUDA GP engine.

    kriging(st, et, ...)      observation / test split, model construction, pickle + .mat
    runRestarts(fname, nres)  hyper-parameter restarts (sharded over ranks when distributed)
    predict(filename, ...)    gridded prediction, one time slice at a time, NetCDF output
    predictTest / getRMSE     hold-out evaluation
    getGrid, rmse, getData, boundData

Host preprocessing follows the reference (krig.py:42-86,259-381,648-678) in Python 3; the
numerical core (covariance, Cholesky, likelihood, prediction) is ``models.GPRegression`` over
libgp2d.  Deviations, all forced by what the reference tree lacks:
  * kernelType 2 / 3 / 4 build myKernel2.divFreeK / curlFreeK / their sum over (t, y, x) with the
    constructor calls of krig.py:396-404; the reference's ``myKernel2`` module is not in its
    repository, so the covariance is the one specified in myKernel2.py / csrc/hsum.cuh (the
    stream-function / potential construction of myKernel.py:39-52 on an anisotropic space-time
    squared exponential); nKernels copies are summed as krig.py:405-407 (at most 8 terms);
  * kernelType 1 (scalar ARD RBF over t,y,x; krig.py:388) builds one scalar model per velocity
    component (``_v.pkl`` / ``_u.pkl``) on the GPU RBF family; ``scikit_prior`` rebuilds the
    optimised kernel through the scikit-learn look-alike (sklearn_like.py) as krig.py:174-194;
  * a vector kernel gives both velocity components from ONE model, stored as
    ``<output>_combined.pkl`` (``_divFree`` / ``_curlFree``), the names krig.py:398-403 uses.
"""
from __future__ import annotations

import os
import pickle
from datetime import datetime

import numpy as np
import scipy.io as sio

from . import dist as gdist
from . import models
from .kern import RBF
from . import myKernel2
from .sklearn_like import GaussianProcessRegressor, kernels
from .printNCFiles import createNC, openNC, writeNC
from .projection import NAD83

lat0 = 28.8
lon0 = -88.6
x_ori, y_ori = NAD83(lon0, lat0)            # krig.py:17-20

_SUFFIX = {2: "_divFree.pkl", 3: "_curlFree.pkl", 4: "_combined.pkl"}


# ---- data --------------------------------------------------------------------------------
def getData(st, et, laser=1, path=None):
    """time [h], lat, lon, v, u as [time, drifter] and the number of valid points per drifter
    (krig.py:42-77).  laser=1: the filtered LASER pickle; otherwise the NetCDF of simulated
    trajectories (read with scipy's NetCDF-3 reader)."""
    if laser == 1:
        with open(path or "Filtered_2016_2_7.pkl", "rb") as f:
            tr = pickle.load(f, encoding="latin1")
        for name in ("lat", "lon", "u", "v"):          # drifter L_0937 (index 238) is discarded
            getattr(tr, name)[:, 238] = np.nan
        time = (tr.time[st:et] - tr.time[0]) / 3600.
        latt, lont, uob, vob = tr.lat[st:et, :], tr.lon[st:et, :], tr.u[st:et, :], tr.v[st:et, :]
        valid = np.sum(~np.isnan(lont) & ~np.isnan(latt), axis=0).astype(float)
        order = np.argsort(valid)[::-1]
    else:
        from scipy.io import netcdf_file
        f = netcdf_file(path or "Simulations/Output.nc", "r", mmap=False)
        latt = np.array(f.variables["lat"][st:et])
        lont = np.array(f.variables["lon"][st:et])
        vob = np.array(f.variables["v"][st:et])
        uob = np.array(f.variables["u"][st:et])
        time = np.array(f.variables["time"][st:et]) - f.variables["time"][0]
        valid = np.zeros(latt.shape[1]) + time.size
        order = np.arange(latt.shape[1])
        f.close()
    return time, latt[:, order], lont[:, order], vob[:, order], uob[:, order], valid[order]


def boundData(var, varlim, lat, lon, v, u):
    """Keep drifters whose initial ``var`` lies in ``varlim`` (krig.py:79-86)."""
    il = np.where((var[0, :] >= varlim[0]) & (var[0, :] <= varlim[1]))[0]
    return lat[:, il], lon[:, il], v[:, il], u[:, il]


def split_observations(tob, yob, xob, latt, lont, vob, uob, sample_step=5, skip=5):
    """Observation / test split of krig.kriging (krig.py:300-369): every |sample_step|-th time
    step and every skip-th drifter are observations, the complement is the test set; points
    with NaN positions are dropped.  Returns two dicts of column vectors."""
    fields = {"t": tob, "y": yob, "x": xob, "lat": latt, "lon": lont, "u": uob, "v": vob}
    nt, nd = tob.shape
    if (sample_step < 0) or (skip > 1):
        ss = abs(sample_step)
        samples = np.arange(0, nt, ss)
        if skip > 1:
            testt = np.arange(nt)
            testd = np.array(sorted(set(range(nd)) - set(range(0, nd, skip))), dtype=int)
        else:
            testd = np.arange(nd)
            testt = np.array(sorted(set(range(nt)) - set(samples)), dtype=int) if ss > 1 else samples
        obs = {k: np.reshape(a[samples, ::skip], [-1, 1]) for k, a in fields.items()}
        tst = {k: np.reshape(a[testt[:, None], testd], [-1, 1]) for k, a in fields.items()}
    else:
        samples = np.arange(0, xob.size, sample_step)
        test = np.array(sorted(set(range(xob.size)) - set(samples)), dtype=int)
        obs = {k: np.reshape(a, [-1])[samples, None] for k, a in fields.items()}
        tst = {k: np.reshape(a, [-1])[test, None] for k, a in fields.items()}
    for d in (obs, tst):
        ok = np.where(~np.isnan(d["x"][:, 0]) & ~np.isnan(d["y"][:, 0]))[0]
        for k in d:
            d[k] = d[k][ok]
    return obs, tst


def make_kernel(kernelType):
    if kernelType == 1:
        return RBF(input_dim=3, ARD=True)                      # krig.py:388
    if kernelType == 2:
        return myKernel2.divFreeK(input_dim=3, active_dims=[0, 1, 2], var=1., lt=1., ly=1., lx=1.)     # krig.py:397
    if kernelType == 3:
        return myKernel2.curlFreeK(input_dim=3, active_dims=[0, 1, 2], var=1., lt=1., ly=1., lx=1.)    # krig.py:401
    if kernelType == 4:
        return myKernel2.divFreeK(input_dim=3) + myKernel2.curlFreeK(input_dim=3)                      # krig.py:404
    raise ValueError("kernelType must be 1 (scalar ARD RBF), 2 (divergence-free), 3 (curl-free) or 4 (both)")


# ---- workflows ---------------------------------------------------------------------------
def kriging(st, et, lalim=[0, 0], lolim=[0, 0], sample_step=5, skip=5, nKernels=1, output='rbfModel',
            pkg='GPy', kernelType=1, laser=1, data=None):
    """Build the (un-optimised) GP model for drifter data between time steps st and et and
    pickle it, like krig.kriging (krig.py:259-418).  ``data`` may carry the tuple getData
    returns (time, lat, lon, v, u, valid) to bypass file I/O."""
    startTime = datetime.now()
    time, latt, lont, vob, uob, _ = data if data is not None else getData(st, et, laser)
    if lolim[1] > lolim[0]:
        latt, lont, vob, uob = boundData(lont, lolim, latt, lont, vob, uob)
    if lalim[1] > lalim[0]:
        latt, lont, vob, uob = boundData(latt, lalim, latt, lont, vob, uob)
    xob, yob = NAD83(lont, latt)
    tob = np.repeat(np.asarray(time)[:, None], latt.shape[1], axis=1)
    bad = np.isnan(lont)
    xob = np.where(bad, np.nan, (xob - x_ori) / 1000.)          # km
    yob = np.where(bad, np.nan, (yob - y_ori) / 1000.)
    o, t = split_observations(tob, yob, xob, latt, lont, vob, uob, sample_step, skip)
    print('number of observations: ' + str(np.size(o["v"])))
    # From here on, always T, Y, X order; velocities stacked [v; u] (krig.py:376-394)
    X = np.concatenate([o["t"], o["y"], o["x"]], axis=1)
    Xt = np.concatenate([t["t"], t["y"], t["x"]], axis=1)
    obs = np.concatenate([o["v"], o["u"]], axis=0)
    obst = np.concatenate([t["v"], t["u"]], axis=0)
    k2 = make_kernel(kernelType)
    if kernelType == 1:
        # scalar kernels: k = k2 + k2 + ... (krig.py:405-407), one model per component, and the
        # .mat keeps the two-column [v, u] layout of krig.py:378-382
        k = k2.copy()
        for _ in range(nKernels - 1):
            k = k + k2
        model = []
        for comp, suffix in ((o["v"], '_v.pkl'), (o["u"], '_u.pkl')):
            mdl = models.GPRegression(X, comp, k.copy())
            mdl.pickle(output + suffix)
            model.append(mdl)
        obs = np.concatenate([o["v"], o["u"]], axis=1)
        obst = np.concatenate([t["v"], t["u"]], axis=1)
    else:
        k = k2.copy()
        for _ in range(nKernels - 1):                           # krig.py:405-407
            k = k + k2
        model = models.GPRegression(X, obs, k)
        model.pickle(output + _SUFFIX[kernelType])
    sio.savemat(output + '.mat', {'Xo': X, 'obs': obs, 'Xt': Xt,
                                  'LL_o': np.concatenate([o["t"], o["lat"], o["lon"]], axis=1),
                                  'LL_t': np.concatenate([t["t"], t["lat"], t["lon"]], axis=1),
                                  'test_points': obst})
    print('End of script, time : ' + str(datetime.now() - startTime))
    return model


def _model_file(fname):
    for suf in ("",) + tuple(_SUFFIX.values()):
        p = fname + suf if suf else fname + '.pkl'
        if os.path.isfile(p):
            return p
    raise IOError("no model pickle found for '%s'" % fname)


def runRestarts(fname, nres=10, nKernels=2, seed=None, max_iters=1000):
    """Hyper-parameter fit with ``nres`` restarts (krig.py:430-468).  Under torch.distributed
    the restarts are sharded round-robin over the ranks and the best run is gathered.  For the
    scalar models of kernelType=1 pass the per-component name (``<output>_v`` / ``<output>_u``)."""
    startTime = datetime.now()
    path = _model_file(fname)
    model = models.load(path)
    hyp_old = model.param_array.copy()
    rank, world = gdist.world()
    model.optimize_restarts(messages=False, verbose=False, num_restarts=nres, seed=seed, max_iters=max_iters,
                            rank=rank, world=world, robust=True)
    best = gdist.gather_best(model)
    hyp = model.param_array
    if not np.isfinite(best):
        print('runRestarts: every restart failed; %s left as it was' % path)
        return model
    if rank == 0:
        model.pickle(path)
        print('Optimized Hyperparameters =================================================')
        for name, a, b in zip(model.parameter_names(), hyp_old, hyp):
            print('%-28s = %s   :   %s' % (name, a, b))
        print('===========================================================================')
        print('End of script, time : ' + str(datetime.now() - startTime))
    return model


def getGrid(to, yo, xo, dt=0.5, dx=0.5, xL=40, yL=40):
    """Regular (t, y, x) grid around the observations, rows ordered time-major then y then x
    (krig.py:648-678).  Returns Xgrid[nt*ny*nx, 3], tg, yg, xg."""
    def span(v, L):
        if (np.max(v) - np.min(v)) > L:
            return np.mean(v) - L / 2, np.mean(v) + L / 2
        return np.min(v) - dx, np.max(v) + dx
    xmin, xmax = span(xo, xL)
    ymin, ymax = span(yo, yL)
    xg = np.arange(xmin, xmax, dx)
    yg = np.arange(ymin, ymax, dx)
    tg = np.arange(np.min(to), np.max(to), dt)
    Yg, Tg, Xg = np.meshgrid(yg, tg, xg)
    return (np.concatenate([np.reshape(Tg, [Tg.size, 1]), np.reshape(Yg, [Yg.size, 1]),
                            np.reshape(Xg, [Xg.size, 1])], axis=1), tg, yg, xg)


def predict(filename, tlim=[0, 0], ylim=[0, 0], xlim=[0, 0], dt=0.5, dx=0.5, xL=40, yL=40, Simul=0,
            write=True):
    """Gridded posterior mean and variance, one time slice at a time (krig.py:471-574).  Under
    torch.distributed each slice's grid points are sharded over the ranks."""
    startTime = datetime.now()
    if os.path.isfile(filename + '_v.pkl') and os.path.isfile(filename + '_u.pkl'):
        return _predict_scalar(filename, tlim, ylim, xlim, dt, dx, xL, yL, write, startTime)
    model = models.load(_model_file(filename))
    hyp = model.param_array
    if (ylim[0] == ylim[1]) and (xlim[0] == xlim[1]):
        Xo = sio.loadmat(filename + '.mat')['Xo']
        if Simul == 1:
            raise NotImplementedError("the NCOM grid branch needs osprein_2013_8.nc (absent upstream)")
        Xp, tp, yp, xp = getGrid(Xo[:, 0], Xo[:, 1], Xo[:, 2])
    else:
        Xp, tp, yp, xp = getGrid(tlim, ylim, xlim, dt, dx, xL, yL)
    inc = yp.size * xp.size
    V, U, VVar, UVar = [], [], [], []
    for i in range(tp.size):
        Xp2 = Xp[i * inc:(i + 1) * inc]                       # rows (t, y, x) of this time slice (krig.py:541-543)
        if not (model.hsum or model.spacetime):
            Xp2 = Xp2[:, 1:3]                                 # purely spatial kernels see (y, x)
        mean, var = gdist.predict_sharded(model._gp, Xp2, include_noise=True)     # GPy adds the noise
        mean, var = mean.cpu().numpy(), var.cpu().numpy()
        V.append(mean[:inc]); U.append(mean[inc:]); VVar.append(var[:inc]); UVar.append(var[inc:])
    shape = [tp.size, yp.size, xp.size]
    V, U, VVar, UVar = (np.reshape(np.concatenate(a), shape) for a in (V, U, VVar, UVar))
    if write and gdist.world()[0] == 0:
        createNC(filename + '.nc', tp, yp, xp, hyp)
        fi = openNC(filename + '.nc', 'a')
        for name, arr in (('v', V), ('u', U), ('vvar', VVar), ('uvar', UVar),
                          ('hyperparam_v', hyp), ('hyperparam_u', hyp)):
            writeNC(fi, name, arr)
        fi.close()
        print('End of script, time : ' + str(datetime.now() - startTime))
    return Xp, V, U, VVar, UVar


def _predict_scalar(filename, tlim, ylim, xlim, dt, dx, xL, yL, write, startTime):
    """krig.predict for the scalar models: model_v and model_u predicted slice by slice
    (krig.py:478-483,541-557)."""
    model_v = models.load(filename + '_v.pkl')
    model_u = models.load(filename + '_u.pkl')
    if (ylim[0] == ylim[1]) and (xlim[0] == xlim[1]):
        Xo = sio.loadmat(filename + '.mat')['Xo']
        Xp, tp, yp, xp = getGrid(Xo[:, 0], Xo[:, 1], Xo[:, 2])
    else:
        Xp, tp, yp, xp = getGrid(tlim, ylim, xlim, dt, dx, xL, yL)
    inc = yp.size * xp.size
    shape = [tp.size, yp.size, xp.size]
    out = {}
    for name, mdl in (('v', model_v), ('u', model_u)):
        mean, var = [], []
        for i in range(tp.size):
            mu, vv = mdl.predict(Xp[i * inc:(i + 1) * inc])
            mean.append(mu[:, 0]); var.append(vv[:, 0])
        out[name] = np.reshape(np.concatenate(mean), shape)
        out[name + 'var'] = np.reshape(np.concatenate(var), shape)
    if write and gdist.world()[0] == 0:
        createNC(filename + '.nc', tp, yp, xp, model_v.param_array)
        fi = openNC(filename + '.nc', 'a')
        for name in ('v', 'u', 'vvar', 'uvar'):
            writeNC(fi, name, out[name])
        writeNC(fi, 'hyperparam_v', model_v.param_array)
        writeNC(fi, 'hyperparam_u', model_u.param_array)
        fi.close()
        print('End of script, time : ' + str(datetime.now() - startTime))
    return Xp, out['v'], out['u'], out['vvar'], out['uvar']


def _predict_points(filename, Xt):
    """Mean / variance of (v, u) at arbitrary (t, y, x) rows from whichever models ``filename``
    names: one vector-valued model or the scalar pair ``_v.pkl`` / ``_u.pkl``."""
    Nt = Xt.shape[0]
    if os.path.isfile(filename + '_v.pkl') and os.path.isfile(filename + '_u.pkl'):
        mv, vv = models.load(filename + '_v.pkl').predict(Xt)
        mu, vu = models.load(filename + '_u.pkl').predict(Xt)
        return mv, vv, mu, vu
    mean, var = models.load(_model_file(filename)).predict(Xt)
    return mean[:Nt], var[:Nt], mean[Nt:], var[Nt:]


def _test_columns(obst, Nt):
    """Test observations as (v, u) columns: stacked [v; u] for the vector models, two columns for
    the scalar ones (krig.py:378-394)."""
    obst = np.asarray(obst)
    if obst.shape[1] == 2:
        return obst[:, 0:1], obst[:, 1:2]
    return obst[:Nt], obst[Nt:]


def predictTest(filename):
    """Predict at the held-out test points and save them (krig.py:578-616)."""
    f = sio.loadmat(filename + '.mat')
    Xt, obst = f['Xt'], f['test_points']
    Vp, VpVar, Up, UpVar = _predict_points(filename, Xt)
    out = {'Xt': Xt, 'Vp': Vp, 'VpVar': VpVar, 'Up': Up, 'UpVar': UpVar, 'test_points': obst}
    sio.savemat(filename + '_test.mat', out)
    return out


def testModel1D(model, Xt, test_points):
    """RMSE of a scalar model at the test points (krig.py:637-639)."""
    f, fVar = model.predict(Xt)
    return rmse(f, test_points)


def rmse(ys, y):
    """Root-mean-square error (krig.py:641-645)."""
    error = np.reshape(np.asarray(ys) - np.asarray(y), [-1])
    return np.sqrt(np.mean(np.square(error)))


def getRMSE(filename):
    """RMSE of the v and u predictions at the test points (krig.py:620-635)."""
    f = sio.loadmat(filename + '.mat')
    Xt, obst = f['Xt'], f['test_points']
    Vp, _, Up, _ = _predict_points(filename, Xt)
    vt, ut = _test_columns(obst, Xt.shape[0])
    rmse_v, rmse_u = rmse(Vp, vt), rmse(Up, ut)
    print(rmse_v)
    print(rmse_u)
    return rmse_v, rmse_u


def scikit_prior(filename0, varname='v', dt=0, tlim=6, radar='', xlim=[0, 0], ylim=[0, 0], dx=0, ind=0, xrange=3):
    """Prediction with the optimised hyper-parameters through the scikit-learn formulation
    (krig.py:88-207): window the observations around the target time / region, rebuild
    ``HP[0]*RBF([..]) (+ HP[4]*RBF([..])) + WhiteKernel(noise)`` from the pickled model's
    ``param_array``, fit with fixed hyper-parameters, predict mean and standard deviation on the
    grid and append them to a NetCDF file.  Returns (outFile, U, Uvar)."""
    startTime = datetime.now()
    if "res" in filename0:                                  # <dir>res<k>/<name>: data live in <dir>
        dir0, a = filename0.split("res")
        _, fname0 = a.split("/")
        fname0 = dir0 + fname0
    else:
        fname0 = filename0
    fm = sio.loadmat(fname0 + '.mat')
    print('Longitude limits:', xlim)
    print('Latitude limits :', ylim)
    if radar != '':
        raise NotImplementedError("the radar-grid branch needs the radar NetCDF files (absent upstream)")
    if (xlim[1] > xlim[0]) and (ylim[1] > ylim[0]):
        X, tcenter, yg, xg = getGrid([dt, dt + 1], ylim, xlim, 1, dx)
        filename = filename0 + '_cyc'
    else:                                                    # pre-existing grid (NetCDF written by predict)
        from scipy.io import netcdf_file
        f = netcdf_file(filename0 + '.nc', 'r', mmap=False)
        xg, yg, tg = np.array(f.variables['x'][:]), np.array(f.variables['y'][:]), np.array(f.variables['time'][:])
        f.close()
        it = dt
        tcenter = np.array([tg[it]])
        Yg, Tg, Xg = np.meshgrid(yg, tg, xg)
        X = np.concatenate([np.reshape(Tg, [Tg.size, 1]), np.reshape(Yg, [Yg.size, 1]), np.reshape(Xg, [Xg.size, 1])], axis=1)
        filename = filename0
        inc = yg.size * xg.size
        X = X[inc * it:inc * it + inc, :]
    filename = filename + '_' + str(np.round(tcenter[0], decimals=2)) + 'h_scikit_'
    outFile = filename + str(ind) + '.nc'
    # observations inside the time / longitude window (krig.py:146-157)
    to, tt = fm['Xo'][:, 0], fm['Xt'][:, 0]
    xo, xt = fm['Xo'][:, 2], fm['Xt'][:, 2]
    ito = np.where((to >= tcenter - tlim) & (to <= tcenter + tlim) & (xo >= xlim[0] - xrange) & (xo <= xlim[1] + xrange))[0]
    itt = np.where((tt >= tcenter - tlim) & (tt <= tcenter + tlim) & (xt >= xlim[0] - xrange) & (xt <= xlim[1] + xrange))[0]
    XT = np.concatenate([fm['Xo'][ito, :], fm['Xt'][itt, :]], axis=0)
    print('Number of observation points: ', np.size(XT, 0))
    obs, obst = fm['obs'][ito, :], fm['test_points'][itt, :]
    model = models.load(fname0 + '_' + varname + '.pkl')
    HP = model.param_array
    col = 1 if varname == 'u' else 0
    u = np.concatenate([obs[:, col], obst[:, col]])[:, None]
    N = HP.size - 1
    noise = HP[-1]
    print('noise = ' + str(HP[-1]))
    k = HP[0] * kernels.RBF(length_scale=[HP[1], HP[2], HP[3]])
    if N > 5:
        k = k + HP[4] * kernels.RBF(length_scale=[HP[5], HP[6], HP[7]])
    k = k + kernels.WhiteKernel(noise_level=noise)
    print(k)
    model_u = GaussianProcessRegressor(kernel=k, optimizer=None)
    model_u.fit(XT, u)
    U, Ustd = model_u.predict(X, return_std=True)
    U = np.reshape(U, [tcenter.size, yg.size, xg.size])
    Ustd = np.reshape(Ustd, [tcenter.size, yg.size, xg.size])
    if not os.path.isfile(outFile):
        createNC(outFile, tcenter, yg, xg, HP)
    fi = openNC(outFile, 'a')
    writeNC(fi, varname, U)
    writeNC(fi, varname + 'var', Ustd ** 2)
    writeNC(fi, 'hyperparam_' + varname, HP)
    fi.close()
    print('End of script, time : ' + str(datetime.now() - startTime))
    return outFile, U, Ustd ** 2



def scikitSnapshot(filename, var, dt, ylim=[0, 0], xlim=[0, 0]):
    """One time slice of the grid in ``filename.nc`` predicted with the pickled scikit-learn-style model
    ``filename_scikit_u.pkl`` (var == 1) or ``filename_scikit_v.pkl`` and written to
    ``<filename>_<time>h_scikit.nc`` (krig.py:210-256).  Upstream the function cannot run: it uses the
    undefined names ``varname`` / ``covarname`` / ``startTime`` (krig.py:253-256) and indexes
    ``ylim[1>ylim[0]]`` (krig.py:218); here the variable is 'u' / 'v' by ``var``, its variance goes to
    '<name>var' as in scikit_prior, and both limits are compared the way xlim is.  The slice is
    ``tg.size // 2 + dt``.  Returns (outFile, U, Uvar)."""
    import pickle
    from scipy.io import netcdf_file
    startTime = datetime.now()
    f = netcdf_file(filename + '.nc', 'r', mmap=False)
    HPV = np.array(f.variables['hyperparam_v'][:])
    xg, yg, tg = np.array(f.variables['x'][:]), np.array(f.variables['y'][:]), np.array(f.variables['time'][:])
    f.close()
    if (xlim[1] > xlim[0]) and (ylim[1] > ylim[0]):
        xg = xg[np.where((xg >= xlim[0]) & (xg <= xlim[1]))]
        yg = yg[np.where((yg >= ylim[0]) & (yg <= ylim[1]))]
    Yg, Tg, Xg = np.meshgrid(yg, tg, xg)
    X = np.concatenate([np.reshape(Tg, [Tg.size, 1]), np.reshape(Yg, [Yg.size, 1]), np.reshape(Xg, [Xg.size, 1])], axis=1)
    inc = yg.size * xg.size
    it = tg.size // 2 + dt
    i2 = inc * it
    outFile = filename + '_' + str(tg[it]) + 'h_scikit.nc'
    tg = np.array([tg[it]])
    X2 = X[i2:i2 + inc, :]
    varname = 'u' if var == 1 else 'v'
    with open(filename + '_scikit_' + varname + '.pkl', 'rb') as inp:
        model_u = pickle.load(inp)
    U, Ustd = model_u.predict(X2, return_std=True)
    U = np.reshape(U, [tg.size, yg.size, xg.size])
    Ustd = np.reshape(Ustd, [tg.size, yg.size, xg.size])
    if not os.path.isfile(outFile):
        createNC(outFile, tg, yg, xg, HPV)
    fi = openNC(outFile, 'a')
    writeNC(fi, varname, U)
    writeNC(fi, varname + 'var', Ustd ** 2)
    fi.close()
    print('End of script, time : ' + str(datetime.now() - startTime))
    return outFile, U, Ustd ** 2
