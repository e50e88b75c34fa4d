"""CPU-side tests of the host logic around the GPU path: parameter transforms, projection,
observation/test split, grid ordering, NetCDF writer, sharding arithmetic and the
torch.distributed gathers (gloo, world_size 2).  No compute call reaches libgp2d here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gp2d_b200 import dist as gdist
from gp2d_b200 import krig, printNCFiles, projection, synthetic
from gp2d_b200.params import Param


# ---- Param (GPy Logexp / Logistic stand-ins; myKernel.py:16-22) ------------------------------
@pytest.mark.parametrize("make,vals", [
    (lambda: Param("l", 1.0).constrain_positive(), [1e-6, 0.3, 2.0, 40.0]),
    (lambda: Param("r", 0.5).constrain_bounded(0, 1), [1e-6, 0.2, 0.5, 0.999]),
])
def test_param_transform_roundtrip_and_derivative(make, vals):
    for v in vals:
        p = make()
        p.value = v
        x = p.to_free()
        p.from_free(x)
        assert p.value == pytest.approx(v, rel=1e-9, abs=1e-12)
        h = 1e-6
        q = make()
        q.from_free(x + h)
        hi = q.value
        q.from_free(x - h)
        lo = q.value
        assert p.dvalue_dfree(x) == pytest.approx((hi - lo) / (2 * h), rel=1e-5, abs=1e-10)


def test_param_transforms_do_not_overflow():
    p = Param("v", 1.0).constrain_positive()
    p.from_free(-754.0)
    assert p.value > 0 and np.isfinite(p.dvalue_dfree(-754.0))
    p.from_free(800.0)
    assert p.value == 800.0 and p.dvalue_dfree(800.0) == 1.0
    r = Param("r", 0.5).constrain_bounded(0, 1)
    r.from_free(-5000.0)
    assert 0 < r.value < 1e-10 and np.isfinite(r.dvalue_dfree(-5000.0))
    r.from_free(5000.0)
    assert r.value <= 1.0 and np.isfinite(r.dvalue_dfree(5000.0))


def test_param_behaves_like_a_float():
    p = Param("length_df", np.array([2.5]))
    assert float(p) == 2.5 and p[0] == 2.5
    assert p * 2 == 5.0 and 2 * p == 5.0 and p ** 2 == 6.25 and 1 / p == 0.4
    p.gradient = 3.0
    assert p.gradient == 3.0
    assert p > 2 and p >= 2.5 and p < 3 and p <= 2.5 and -p == -2.5
    f = Param("x", 1.0).constrain_fixed(4.0)
    f.from_free(9.0)
    assert f.value == 4.0


# ---- projection (krig.py:19-20,291: EPSG:3452, Lambert conformal conic 2SP on GRS80) ---------
def test_projection_origin_and_standard_parallel_scale():
    x, y = projection.NAD83(-(91.0 + 20.0 / 60.0), 28.5)
    assert x == pytest.approx(1000000.0, abs=1e-6) and y == pytest.approx(0.0, abs=1e-6)
    a, f = 6378137.0, 1 / 298.257222101
    e2 = 2 * f - f * f
    for lat in (29.3, 30.7):            # scale factor is exactly 1 on the standard parallels
        dl = 1e-4
        x0, y0 = projection.NAD83(-89.0, lat)
        x1, y1 = projection.NAD83(-89.0 + dl, lat)
        phi = np.radians(lat)
        arc = a * np.cos(phi) / np.sqrt(1 - e2 * np.sin(phi) ** 2) * np.radians(dl)
        assert np.hypot(x1 - x0, y1 - y0) == pytest.approx(arc, rel=1e-7)
    # conformal: a small step north and a small step east are orthogonal and equally scaled
    lat, lon, d = 28.8, -88.6, 1e-4
    x0, y0 = projection.NAD83(lon, lat)
    xe, ye = projection.NAD83(lon + d, lat)
    xn, yn = projection.NAD83(lon, lat + d)
    ve, vn = np.array([xe - x0, ye - y0]), np.array([xn - x0, yn - y0])
    assert abs(ve @ vn) / (np.linalg.norm(ve) * np.linalg.norm(vn)) < 1e-6
    assert np.isnan(projection.NAD83(np.nan, 28.0)[0])


# ---- observation / test split (krig.py:300-369) ----------------------------------------------
def _fields(nt=6, nd=7):
    t = np.repeat(np.arange(nt, dtype=float)[:, None], nd, axis=1)
    base = np.arange(nt * nd, dtype=float).reshape(nt, nd)
    return t, base + 0.1, base + 0.2, base + 0.3, base + 0.4, base + 0.5, base + 0.6


def test_split_time_subsampling():
    t, y, x, la, lo, v, u = _fields()
    o, s = krig.split_observations(t, y, x, la, lo, v, u, sample_step=-2, skip=1)
    np.testing.assert_array_equal(np.unique(o["t"]), [0, 2, 4])
    np.testing.assert_array_equal(np.unique(s["t"]), [1, 3, 5])
    assert o["x"].shape == (21, 1) and s["x"].shape == (21, 1)
    np.testing.assert_allclose(o["u"] - o["x"], 0.4)          # fields stay aligned


def test_split_drifter_skipping_and_nan_filter():
    t, y, x, la, lo, v, u = _fields()
    x[1, 0] = np.nan
    x[2, 1] = np.nan
    o, s = krig.split_observations(t, y, x, la, lo, v, u, sample_step=-1, skip=3)
    # observations: drifters 0, 3, 6 at every time step, minus one NaN
    assert o["x"].shape[0] == 6 * 3 - 1
    # test set: the other 4 drifters at every time step, minus one NaN
    assert s["x"].shape[0] == 6 * 4 - 1
    assert not np.isnan(o["x"]).any() and not np.isnan(s["x"]).any()


def test_split_flat_sampling():
    t, y, x, la, lo, v, u = _fields()
    o, s = krig.split_observations(t, y, x, la, lo, v, u, sample_step=5, skip=1)
    np.testing.assert_array_equal(o["x"][:, 0], np.reshape(x, [-1])[::5])
    assert o["x"].shape[0] + s["x"].shape[0] == x.size


# ---- grid (krig.py:648-678) --------------------------------------------------------------------
def test_getGrid_ordering_time_major_then_y_then_x():
    Xg, tg, yg, xg = krig.getGrid(np.array([0.0, 2.0]), np.array([1.0, 3.0]), np.array([5.0, 6.0]),
                                  dt=0.5, dx=0.5)
    nt, ny, nx = tg.size, yg.size, xg.size
    assert Xg.shape == (nt * ny * nx, 3)
    G = Xg.reshape(nt, ny, nx, 3)
    np.testing.assert_array_equal(G[:, 0, 0, 0], tg)
    np.testing.assert_array_equal(G[0, :, 0, 1], yg)
    np.testing.assert_array_equal(G[0, 0, :, 2], xg)
    # spans wider than the window are centred on the mean (krig.py:652-661)
    Xg2, _, yg2, _ = krig.getGrid(np.array([0.0, 1.0]), np.array([0.0, 100.0]), np.array([0.0, 1.0]), yL=40)
    assert yg2[0] == pytest.approx(30.0) and yg2[-1] < 70.0


def test_rmse():
    assert krig.rmse(np.array([[1.0], [3.0]]), np.array([[0.0], [0.0]])) == pytest.approx(np.sqrt(5.0))


# ---- NetCDF-3 writer (printNCFiles.py:5-44) ------------------------------------------------------
def test_netcdf_layout_roundtrip(tmp_path):
    from scipy.io import netcdf_file
    path = str(tmp_path / "out.nc")
    T, Y, X = np.array([12.0, 13.0]), np.linspace(0, 1, 4), np.linspace(0, 2, 5)
    hyp = np.array([2.0, 3.0, 0.5, 0.05])
    printNCFiles.createNC(path, T, Y, X, hyp)
    fi = printNCFiles.openNC(path, "a")
    data = np.arange(2 * 4 * 5, dtype=float).reshape(2, 4, 5)
    for name in ("v", "u", "vvar", "uvar"):
        printNCFiles.writeNC(fi, name, data)
    printNCFiles.writeNC(fi, "hyperparam_u", hyp)
    printNCFiles.writeNC(fi, "hyperparam_v", hyp)
    fi.close()
    f = netcdf_file(path, "r", mmap=False)
    assert set(f.dimensions) == {"time", "y", "x", "hyperparam"}
    assert f.variables["u"].dimensions == ("time", "y", "x")
    assert f.variables["u"].data.dtype.kind == "f" and f.variables["u"].data.dtype.itemsize == 4
    np.testing.assert_allclose(f.variables["uvar"].data, data)
    np.testing.assert_allclose(f.variables["hyperparam_v"].data, hyp, rtol=1e-6)
    np.testing.assert_allclose(f.variables["time"].data, T)
    f.close()


# ---- synthetic tracks -------------------------------------------------------------
def test_synthetic_snapshot_is_seeded_and_shaped():
    X, y = synthetic.drifter_snapshot(500, config_id=2, seed_offset=3)
    X2, y2 = synthetic.drifter_snapshot(500, config_id=2, seed_offset=3)
    np.testing.assert_array_equal(X, X2)
    np.testing.assert_array_equal(y, y2)
    assert X.shape == (500, 2) and y.shape == (1000,)
    assert np.abs(y).max() < 0.8
    d = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)) + np.eye(500) * 9
    assert d.min() > 0.09                      # lattice jitter keeps points apart
    G = synthetic.prediction_grid(X, 7, 5)
    assert G.shape == (35, 2) and G[1, 0] > G[0, 0] and G[7, 1] > G[0, 1]


# ---- sharding arithmetic --------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 1000, 102400, 1000000])
@pytest.mark.parametrize("ws", [1, 2, 3, 8])
def test_shard_cyclic_is_a_partition_in_whole_blocks(n, ws):
    parts = [gdist.shard_cyclic(n, r, ws) for r in range(ws)]
    allidx = np.sort(np.concatenate(parts))
    np.testing.assert_array_equal(allidx, np.arange(n))               # every point exactly once
    for r, p in enumerate(parts):
        assert np.all(np.diff(p) > 0)
        assert np.all((p // 320) % ws == r)                            # whole 320-point blocks, dealt round robin
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 320


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 1000, 102400, 1000000])
@pytest.mark.parametrize("ws", [1, 2, 3, 8])
def test_shard_range_partitions_whole_tiles(n, ws):
    parts = [gdist.shard_range(n, r, ws) for r in range(ws)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    for (a, b), (c, d) in zip(parts, parts[1:]):
        assert b == c and a <= b
    for a, b in parts:
        assert a % 64 == 0 or a == n
    sizes = [b - a for a, b in parts]
    assert max(sizes) - min(sizes) <= 64 + 63
    assert sorted(sum((gdist.round_robin(10, r, ws) for r in range(ws)), [])) == list(range(10))


# ---- gathers under gloo, world_size 2 ----------------------------------------------------------------
class _FakeGP:
    """predict() of a deterministic function of the grid points (stands in for HelmholtzGP)."""
    device = torch.device("cpu")

    def predict(self, Xs, include_noise=False):
        Xs = torch.as_tensor(np.asarray(Xs))
        m0, m1 = Xs[:, 0] * 2.0, Xs[:, 1] - 1.0
        v0, v1 = Xs[:, 0] ** 2, Xs[:, 1] ** 2 + (0.05 if include_noise else 0.0)
        return torch.cat([m0, m1]), torch.cat([v0, v1])


class _FakeFit:
    """Carries a byte buffer like HelmholtzGP.predict_state()."""

    def __init__(self, rank, nbytes):
        self.buf = torch.full((nbytes,), 7 if rank == 0 else 0, dtype=torch.uint8)
        if rank == 0:
            self.buf[::3] = 1
        self.fitted = rank == 0

    def predict_state(self):
        return self.buf


class _Run:
    def __init__(self, f, x):
        self.f_opt, self.x_opt = f, np.asarray(x, float)


class _FakeModel:
    def __init__(self, runs):
        self.optimization_runs = runs
        self._gp = _FakeGP()
        self.loaded = None

    def _free_params(self):
        return [0, 1, 2]

    def _set_free(self, x):
        self.loaded = np.asarray(x, float)

    def parameters_changed(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # variable-length gather
        loc = torch.arange(3 + 2 * rank, dtype=torch.float64) + 100 * rank
        got = gdist.gather_concat(loc)
        # grid-sharded prediction: 900 points in 320-point blocks -> rank 0 gets blocks 0 and 2, rank 1 block 1
        Xs = np.stack([np.linspace(0, 1, 900), np.linspace(2, 3, 900)], axis=1)
        mean, var = gdist.predict_sharded(_FakeGP(), Xs, include_noise=True)
        # restart sharding: each rank holds its own runs, the best one wins everywhere
        runs = [_Run(5.0 - rank * 3.0 + i, [rank, i, 7.0]) for i in range(2)]
        model = _FakeModel(runs)
        best = gdist.gather_best(model)
        # one factorisation, shipped to the other rank in pieces
        fitst = _FakeFit(rank, 1000)
        gdist.broadcast_fit(fitst, src=0, chunk_bytes=384)
        # drifters sharded over ranks (laser_io_methods.interp_kriging): the per-drifter solver is stubbed,
        # the sharding, the object gather and the assembly are the real ones
        from gp2d_b200 import laser_io_methods as lio
        calls = []

        def fake_kriging(dr, time, t_origin=None, optimize=True, max_iters=200, device=None):
            calls.append(dr.id)
            nT = time.size
            base = float(dr.id[2:])
            return {"lon": np.full(nT, base), "lat": np.full(nT, -base), "varLon": np.full(nT, 0.1), "varLat": np.full(nT, 0.2),
                    "u": np.full(nT - 1, base), "v": np.full(nT - 1, base), "n_samples": np.zeros(nT - 1), "data_freq": np.zeros(nT - 1),
                    "drog_stat": np.ones(nT), "lastDrogTime": -1, "lenLon": base, "lenLat": base, "varianceLon": 1.0,
                    "varianceLat": 1.0, "noiseLon": 1e-7, "noiseLat": 1e-7}
        lio.kriging = fake_kriging
        fleet = [lio.drifter("L_%d" % i, None, np.arange(0.0, 7200.0, 300.0) + 1000.0, np.zeros(24), np.zeros(24)) for i in range(5)]
        tr = lio.interp_kriging(fleet, dt=900, period=0.05, optimize=False, parallel=1)
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), got=got.numpy(), mean=mean.numpy(), var=var.numpy(),
                 best=best, loaded=model.loaded, shard=np.array(gdist.shard_range(200, rank, world)),
                 state=fitst.buf.numpy(), fitted=fitst.fitted, track_lon=tr.lon, track_len=tr.lenLon,
                 track_calls=np.array([int(c[2:]) for c in calls]))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gathers(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(str(tmp_path / ("r%d.npz" % k))) for k in range(2)]
    Xs = np.stack([np.linspace(0, 1, 900), np.linspace(2, 3, 900)], axis=1)
    m_ref, v_ref = _FakeGP().predict(Xs, include_noise=True)
    for k in range(2):
        np.testing.assert_array_equal(r[k]["got"], np.concatenate([np.arange(3.0), np.arange(5.0) + 100]))
        np.testing.assert_array_equal(r[k]["mean"], m_ref.numpy())      # same on every rank, in grid order
        np.testing.assert_array_equal(r[k]["var"], v_ref.numpy())
        assert float(r[k]["best"]) == 2.0                                # rank 1's first run
        np.testing.assert_array_equal(r[k]["loaded"], [1.0, 0.0, 7.0])
    np.testing.assert_array_equal(r[1]["state"], r[0]["state"])
    assert r[0]["state"][0] == 1 and r[0]["state"][1] == 7 and bool(r[1]["fitted"])
    np.testing.assert_array_equal(r[0]["shard"], [0, 128])
    np.testing.assert_array_equal(r[1]["shard"], [128, 200])
    # drifter n ran on rank n % 2 only, and both ranks hold the assembled fleet
    np.testing.assert_array_equal(r[0]["track_calls"], [0, 2, 4])
    np.testing.assert_array_equal(r[1]["track_calls"], [1, 3])
    for k in range(2):
        np.testing.assert_array_equal(r[k]["track_len"], np.arange(5.0))
        np.testing.assert_array_equal(r[k]["track_lon"][:, 0], np.arange(5.0))


def test_param_fix_and_unfix_restore_the_constraint():
    p = Param("ratio", 0.3).constrain_bounded(0, 1)
    p.fix(0.5)
    assert p.constraint == "fixed" and float(p) == 0.5
    p.fix()                                  # fixing twice does not forget the original constraint
    p.unfix()
    assert p.constraint == ("bounded", 0.0, 1.0)
    q = Param("l", 2.0).constrain_positive().constrain_fixed()
    assert q.unconstrain_fixed().constraint == "positive"


def test_i8_issuer_protocol_model():
    """The barrier protocol of predict_i8_kernel (one producer, two MMA issuers in turn, epilogue; csrc/predict_i8.cu)
    in a discrete-event model with one-bit mbarrier parities, out-of-order copy completion and random scheduling:
    no deadlock, no stale stage header, no ring slot written over before its products have completed, no product
    outside its accumulation segment, for both ring sizes; and the model does show the aliasing that indexing the full
    barriers by slot alone would have on the odd ring."""
    import importlib.util
    import random
    spec = importlib.util.spec_from_file_location("i8_protocol_sim", os.path.join(os.path.dirname(__file__), "tools", "i8_protocol_sim.py"))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    for seed in range(120):
        r = random.Random(seed)
        segs = [r.choice([1, 1, 2, 2, 3, 4, 7, 12]) for _ in range(r.randint(1, 8))]
        w = {k: r.choice([0.05, 0.3, 1, 4]) for k in ("producer", "copies", "issuer0", "issuer1", "epilogue", "tensor")}
        for stages in (5, 4):
            assert sim.run(seed, segs, stages=stages, weights=w) == "ok", (seed, segs, stages)
    broken = sum(sim.run(seed, [1, 2, 3, 7, 2, 3], stages=5, nfull=5) != "ok" for seed in range(150))
    assert broken > 0
    # ... and the one a lane-parallel producer has without its emission window (a lane several ring turns ahead reads
    # "free" off a barrier whose parity it cannot interpret)
    assert all(sim.run(seed, [2, 3, 7, 12, 3, 2], stages=5, chunk=16, window=None) != "ok" for seed in range(20))
    assert all(sim.run(seed, [2, 3, 7, 12, 3, 2], stages=5, chunk=16) == "ok" for seed in range(20))
