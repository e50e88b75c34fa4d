"""Per-drifter track interpolation with 1-D Gaussian processes, the step that feeds the kriging
path (laser_io_methods.interp_kriging / interp_kriging2 / kriging, laser_io_methods.py:410-700):
every drifter's raw (time, lon, lat) fixes are interpolated onto a regular time axis by two scalar
GPs (longitude and latitude over time, ``RBF(input_dim=1, variance=1159.68, lengthscale=4.5)`` with
the noise variance preset to 1.756e-7, laser_io_methods.py:464,496-503), optionally optimised per
drifter, and velocities follow from centred differences of the interpolated positions.

The drifters are independent: they are spread over CUDA streams inside a process (``parallel``),
and over ranks (drifter n on rank n % world) when torch.distributed is initialised -- the joblib
loop of interp_kriging2 (laser_io_methods.py:629-630) on GPUs.  Each GP is one
``models.GPRegression`` over the scalar RBF family of libgp2d (gp2d_rbf_*).

Raw-data ingest (readTracks, readLog, drogue files) stays out of scope: the functions here take
the list of drifter objects those readers produce (or anything with .id, .time, .lon, .lat,
.date_time, .drogueLoss, .drogueStat, .launchType).
"""
from __future__ import annotations

import pickle

import numpy as np

from . import dist as gdist
from . import models
from .kern import RBF

# hyper-parameters the reference starts every drifter from (laser_io_methods.py:464,497)
TRACK_VARIANCE = 1159.68
TRACK_LENGTHSCALE = 4.5
TRACK_NOISE = 1.75598244486e-07


class drifter(object):
    """Raw track of one drifter (laser_class.py:15-24)."""

    def __init__(self, drifter_id, date_time, time, lat, lon, droLossDate=None, droStat=1, launchType=1):
        self.id = drifter_id
        self.date_time = date_time
        self.time = np.asarray(time, dtype=np.float64)
        self.lat = np.asarray(lat, dtype=np.float64)
        self.lon = np.asarray(lon, dtype=np.float64)
        self.drogueLoss = droLossDate
        self.drogueStat = droStat
        self.launchType = launchType


class interpolated_tracks(object):
    """Result container with the attribute names of laser_class.interpolated_tracks
    (laser_class.py:27-55), which GP_laser / krig.getData read back."""

    def __init__(self, drifter_id, time, lon, lat, u, v, date0, time0, nsamples, mdt, drog_stat0, drog_stat,
                 lDrogueTime, dLossDate, launchType, varLon=0, varLat=0, lenLon=0, lenLat=0, varianceLon=0,
                 varianceLat=0, noiseLon=0, noiseLat=0):
        self.id, self.time, self.date0, self.time0 = drifter_id, time, date0, time0
        self.lat, self.lon = lat, lon
        if np.size(varLon) > 1:
            self.pos_varLon, self.pos_varLat = varLon, varLat
            self.lenLon, self.lenLat = lenLon, lenLat
            self.varianceLon, self.varianceLat = varianceLon, varianceLat
            self.noiseLon, self.noiseLat = noiseLon, noiseLat
        self.u, self.v = u, v
        self.n_samples, self.data_freq = nsamples, mdt
        self.drogueStat0, self.drogueStat = drog_stat0, drog_stat
        self.lastDrogueTime, self.lossDate, self.launchType = lDrogueTime, dLossDate, launchType


def save_object(obj, filename):
    with open(filename, "wb") as f:                        # laser_io_methods.py:23-30
        pickle.dump(obj, f, protocol=2)


def read_object(filename):
    with open(filename, "rb") as f:                        # laser_io_methods.py:32-41
        return pickle.load(f, encoding="latin1")


def countDataPoints(time, tdata):
    """Per interval [time[i], time[i+1]): number of raw fixes and their mean spacing
    (laser_io_methods.py:328-335; NaN where an interval holds fewer than two fixes, as np.mean of
    an empty slice gives there)."""
    time = np.asarray(time, dtype=np.float64)
    tdata = np.asarray(tdata, dtype=np.float64)
    n = time.size - 1
    N, dt = np.zeros(max(n, 0)), np.full(max(n, 0), np.nan)
    for i in range(n):
        sel = tdata[(tdata >= time[i]) & (tdata < time[i + 1])]
        N[i] = sel.size
        if sel.size > 1:
            dt[i] = np.mean(np.diff(sel))
    return N, dt


def _window(dr, time):
    """Grid steps covered by the drifter and the raw fixes that bracket them
    (laser_io_methods.py:468-486): the fixes inside [time[it][0], time[it][-1]] plus one on each
    side when the track extends beyond."""
    t = np.asarray(dr.time, dtype=np.float64)
    it = np.where((time >= t[0]) & (time <= t[-1]))[0]
    if it.size == 0:
        return it, np.zeros(0, dtype=int)
    inside = np.where((t >= time[it][0]) & (t <= time[it][-1]))[0]
    if inside.size == 0:
        # no fix between the first and last covered step: bracket with the neighbours
        lo = int(np.searchsorted(t, time[it][0], side="right")) - 1
        sel = np.arange(max(lo, 0), min(lo + 2, t.size))
        return it, sel
    lo, hi = inside[0], inside[-1]
    if lo > 0:
        lo -= 1
    if t[-1] > t[hi]:
        hi += 1
    return it, np.arange(lo, hi + 1)


def _track_model(X, Y, optimize, max_iters, device, start=None):
    """One scalar track model.  ``start`` = (variance, lengthscale, noise) to begin the optimisation from;
    None = the preset values.  Upstream builds ONE kernel object before the drifter loop and reuses it for
    every drifter and both coordinates (laser_io_methods.py:464-503), so there each optimisation begins at
    the previous model's optimum; the default here starts every model from the presets, which makes the
    drifters independent (they run concurrently and on different ranks) but can end in a different
    L-BFGS-B optimum than the upstream chain.  ``warm_start=True`` in kriging() / interp_kriging() passes
    the previous optimum along and reproduces the upstream order of starts (sequential only)."""
    var0, len0, noise0 = (TRACK_VARIANCE, TRACK_LENGTHSCALE, TRACK_NOISE) if start is None else start
    k = RBF(input_dim=1, variance=var0, lengthscale=len0)
    m = models.GPRegression(X, Y, k, device=device)
    m.Gaussian_noise = noise0
    if optimize:
        m.optimize(max_iters=max_iters, messages=False)
    return m


def kriging(dr, time, t_origin=None, optimize=True, max_iters=200, device=None, start=None):
    """One drifter (laser_io_methods.py:637-700): dict with the interpolated lon / lat and their
    posterior variances on ``time`` (NaN outside the drifter's life span), centred-difference
    velocities, the per-interval data counts and the hyper-parameters of both models.  Times are
    seconds; the GP input is hours since ``t_origin`` (default time[0]).  Mean and variance come from
    the final model (interp_kriging upstream keeps the variance from before the optimisation,
    laser_io_methods.py:499,520; its joblib twin uses the optimised one, :666-668, as here).
    ``start``: (variance, lengthscale, noise) the first (longitude) model starts from; the latitude model
    then starts from the longitude optimum and the last optimum is returned under "next_start" -- the
    chain of starts of the upstream loop, which shares one kernel object (see _track_model)."""
    time = np.asarray(time, dtype=np.float64)
    t0 = time[0] if t_origin is None else float(t_origin)
    nT = time.size
    out = {"lon": np.full(nT, np.nan), "lat": np.full(nT, np.nan), "varLon": np.full(nT, np.nan),
           "varLat": np.full(nT, np.nan), "u": np.full(nT - 1, np.nan), "v": np.full(nT - 1, np.nan),
           "n_samples": np.full(nT - 1, np.nan), "data_freq": np.full(nT - 1, np.nan), "drog_stat": np.zeros(nT),
           "lenLon": np.nan, "lenLat": np.nan, "varianceLon": np.nan, "varianceLat": np.nan,
           "noiseLon": np.nan, "noiseLat": np.nan}
    it, sel = _window(dr, time)
    if it.size and sel.size:
        X = ((np.asarray(dr.time, dtype=np.float64)[sel] - t0) / 3600.)[:, None]
        Tg = ((time[it] - t0) / 3600.)[:, None]
        for name, raw in (("Lon", dr.lon), ("Lat", dr.lat)):
            Y = np.asarray(raw, dtype=np.float64)[sel][:, None]
            m = _track_model(X, Y, optimize, max_iters, device, start=start)
            if start is not None:
                start = (m.rbf.variance[0], m.rbf.lengthscale[0], m.Gaussian_noise[0])
            # cond(K) ~ n variance / noise ~ 1e10-1e13 for these models: predict() takes the iterated
            # solve by itself, not the fused explicit-inverse path (engine.refined_predict)
            mean, var = m.predict(Tg)
            out[name.lower()][it] = mean[:, 0]
            out["var" + name][it] = var[:, 0]
            out["len" + name] = m.rbf.lengthscale[0]
            out["variance" + name] = m.rbf.variance[0]
            out["noise" + name] = m.Gaussian_noise[0]
        if it.size > 1:
            # velocities between consecutive grid steps, metres per second (laser_io_methods.py:541-545)
            dts = np.diff(time[it])
            latm = 0.5 * (out["lat"][it][1:] + out["lat"][it][:-1])
            out["u"][it[1:] - 1] = np.diff(out["lon"][it]) * 111000. * np.cos(latm * np.pi / 180.) / dts
            out["v"][it[1:] - 1] = np.diff(out["lat"][it]) * 111000. / dts
            M1, M2 = countDataPoints(time[it], dr.time)
            out["n_samples"][it[1:] - 1] = M1
            out["data_freq"][it[1:] - 1] = M2
    # drogue status along the time axis (laser_io_methods.py:528-538)
    if start is not None:
        out["next_start"] = start
    loss, dates = getattr(dr, "drogueLoss", None), getattr(dr, "date_time", None)
    before = [] if loss is None or dates is None else [i for i, d in enumerate(dates) if d < loss]
    if before:
        out["lastDrogTime"] = float(np.asarray(dr.time)[max(before)])
        out["drog_stat"][time <= out["lastDrogTime"]] = 1
    else:
        out["lastDrogTime"] = -1
        out["drog_stat"][:] = 1
    return out


def interp_kriging(data, dt=900, period=10, optimize=True, max_iters=200, output=None, parallel=4, device=None,
                   warm_start=False):
    """All drifters (laser_io_methods.py:410-570 and its joblib twin :576-633): ``data`` is the list
    of raw drifter objects, data[0] the first one released.  Returns an ``interpolated_tracks`` and
    pickles it to ``output`` when given.  ``parallel`` drifters are in flight at a time on separate
    CUDA streams; under torch.distributed drifter n runs on rank n % world and rank 0 assembles (and
    writes) the result, which every rank returns.  ``warm_start=True`` reproduces the upstream chain of
    optimisation starts (one shared kernel object, laser_io_methods.py:464-503): every model begins at the
    optimum of the one before it, which makes the drifters sequential (parallel and the rank sharding are
    then switched off)."""
    import torch
    t_first = float(np.asarray(data[0].time)[0])
    time = np.arange(float(np.asarray(data[0].time)[1]), t_first + period * 86400., dt)
    time2 = ((time - t_first) / 3600.)[:, None]                      # hours, [T,1] (laser_io_methods.py:428-429)
    rank, world = gdist.world()
    mine = [n for n in range(len(data)) if n % world == rank]
    results = {}

    def work(todo):
        for n in todo:
            results[n] = kriging(data[n], time, t_origin=t_first, optimize=optimize, max_iters=max_iters, device=device)

    parallel = max(1, min(int(parallel), len(mine)))
    if warm_start:
        start = (TRACK_VARIANCE, TRACK_LENGTHSCALE, TRACK_NOISE)
        for n in range(len(data)):                                  # every rank runs the whole chain
            results[n] = kriging(data[n], time, t_origin=t_first, optimize=optimize, max_iters=max_iters, device=device,
                                 start=start)
            start = results[n].pop("next_start", start)
        world = 1
    elif parallel == 1:
        work(mine)
    else:
        import queue
        import threading
        pending, errors = queue.SimpleQueue(), []
        for n in mine:
            pending.put(n)

        def pull():
            while True:
                try:
                    yield pending.get_nowait()
                except queue.Empty:
                    return

        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())

        def run():
            try:
                with torch.cuda.device(dev), torch.cuda.stream(torch.cuda.Stream(dev)):
                    work(pull())
                    torch.cuda.current_stream().synchronize()
            except Exception as e:                                   # re-raised in the caller's thread
                errors.append(e)
        threads = [threading.Thread(target=run) for _ in range(parallel)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
    if world > 1:
        import torch.distributed as tdist
        parts = [None] * world
        tdist.all_gather_object(parts, results)
        results = {n: r for part in parts for n, r in part.items()}
    N, T = len(data), time.size
    stack = lambda key: np.stack([results[n][key] for n in range(N)]) if N else np.zeros((0, T))
    vec = lambda key: np.array([results[n][key] for n in range(N)])
    tracks = interpolated_tracks(
        [d.id for d in data], time2, stack("lon"), stack("lat"), stack("u"), stack("v"),
        [d.date_time[0] if getattr(d, "date_time", None) is not None and len(d.date_time) else None for d in data],
        [float(np.asarray(d.time)[0]) for d in data], stack("n_samples"), stack("data_freq"),
        np.array([getattr(d, "drogueStat", 1) for d in data], dtype=float), stack("drog_stat"), vec("lastDrogTime"),
        [getattr(d, "drogueLoss", None) for d in data], np.array([getattr(d, "launchType", 1) for d in data], dtype=float),
        stack("varLon"), stack("varLat"), vec("lenLon"), vec("lenLat"), vec("varianceLon"), vec("varianceLat"),
        vec("noiseLon"), vec("noiseLat"))
    if output is not None and rank == 0:
        save_object(tracks, output)
    return tracks


interp_kriging2 = interp_kriging          # the joblib variant upstream (laser_io_methods.py:576-633): same result
