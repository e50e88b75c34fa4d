"""Parity of the int8-sliced (tcgen05 / TMEM) predictive kernel, csrc/predict_i8.cu, against the oracle
(GP_laser.py:113-140 restated in oracle/gp_oracle.py) and against the fp64 tensor-pipe kernel it replaces on the
Helmholtz families.  Tolerances are BASELINE.json's: 1e-8 relative on mean and variance.  The kernel is selected
per fit (gp2d_set_option / engine.set_predict_i8): 6 or 7 base-256 digit slices, or the fp64 kernel.
Run on the B200 box:  pytest tests -m gpu"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():          # collected on the CPU box, run on the GPU box
    pytest.skip("no CUDA device", allow_module_level=True)

import gp2d_b200 as gp                      # noqa: E402
from gp2d_b200 import synthetic             # noqa: E402
from oracle import gp_oracle as orc        # noqa: E402

THETA, NOISE = (1.3, 3.1, 0.2), 0.05


@pytest.fixture(autouse=True)
def _auto_mode_afterwards():
    yield
    gp.set_predict_i8(0)


def _gate(m):
    """Slice count the fit chose (an int in the last 256-byte block of the fit workspace, after info)."""
    return int(m.ws[-256:].view(torch.int32)[2].item())


def _fit_predict(mode, X, y, Xs, theta=THETA, noise=NOISE, cls=None, **kw):
    gp.set_predict_i8(mode)
    m = (cls or gp.HelmholtzGP)(X, y, *theta, noise, **kw)
    m.fit()
    mean, var = m.predict(Xs)
    return m, mean.cpu().numpy(), var.cpu().numpy()


@pytest.mark.parametrize("mode", [6, 7])
@pytest.mark.parametrize("N,side", [(700, 45), (1000, 33)])
def test_i8_predict_vs_oracle(mode, N, side):
    X, y = synthetic.drifter_snapshot(N, config_id=2, seed_offset=N)
    Xs = synthetic.prediction_grid(X, side, side)
    m, mean, var = _fit_predict(mode, X, y, Xs)
    assert _gate(m) == mode
    f = orc.fit(X, y, *THETA, NOISE)
    mo, vo = orc.predict(X, f, *THETA, Xs)
    np.testing.assert_allclose(mean, mo, rtol=1e-8, atol=1e-9 * np.abs(mo).max())
    np.testing.assert_allclose(var, vo, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("N", [1, 5, 63, 64, 65, 130, 449])
@pytest.mark.parametrize("M", [1, 39, 40, 41, 81, 1300])
def test_i8_ragged_sizes_vs_fp64_kernel(N, M):
    """Every padding case of the observation rows (128-row blocks, 16-observation k-steps) and of the column
    tiles (40 grid points with 6 slices, 32 with 7)."""
    rng = np.random.default_rng(100 * N + M)
    X, y = synthetic.drifter_snapshot(N, config_id=2, seed_offset=7)
    Xs = np.stack([rng.uniform(X[:, 0].min() - 1, X[:, 0].max() + 1, M), rng.uniform(X[:, 1].min() - 1, X[:, 1].max() + 1, M)], 1)
    _, m1, v1 = _fit_predict(1, X, y, Xs)
    for mode in (6, 7):
        m, mm, vv = _fit_predict(mode, X, y, Xs)
        assert _gate(m) == mode
        np.testing.assert_allclose(mm, m1, rtol=1e-9, atol=1e-10 * max(np.abs(m1).max(), 1e-300))
        np.testing.assert_allclose(vv, v1, rtol=2e-9, atol=1e-13)


@pytest.mark.parametrize("mode", [6, 7])
def test_i8_partition_invariance_is_bitwise(mode):
    """Integer products are exact and the fp64 recombination has a fixed order: any cut of the grid gives the same
    bits (what dist.predict_sharded and the time-slice loop of krig.py:539-557 rely on)."""
    X, y = synthetic.drifter_snapshot(900, config_id=2, seed_offset=1)
    Xs = synthetic.prediction_grid(X, 60, 50)
    gp.set_predict_i8(mode)
    m = gp.HelmholtzGP(X, y, *THETA, NOISE)
    m.fit()
    M = Xs.shape[0]
    mean, var = m.predict(Xs)
    for lo, hi in [(0, 1), (17, 1234), (1234, M), (M - 41, M)]:
        ms, vs = m.predict(Xs[lo:hi])
        k = hi - lo
        assert torch.equal(ms[:k], mean[lo:hi]) and torch.equal(ms[k:], mean[M + lo:M + hi])
        assert torch.equal(vs[:k], var[lo:hi]) and torch.equal(vs[k:], var[M + lo:M + hi])
    # and from one run to the next
    m2, v2 = m.predict(Xs)
    assert torch.equal(m2, mean) and torch.equal(v2, var)


def test_i8_slice_count_follows_the_conditioning():
    """gp2d_fit picks the slice count from k** / (noise + jitter) sqrt(n / 4000): 6, 7, then the fp64 kernel; the robust
    (ill-conditioned) mode always keeps the fp64 kernel.  Each choice meets the 1e-8 bar against the oracle."""
    X, y = synthetic.drifter_snapshot(600, config_id=2, seed_offset=2)
    Xs = synthetic.prediction_grid(X, 30, 30)
    for noise, want in [(0.05, 6), (0.01, 6), (1e-3, 7), (1e-4, 7), (1e-5, 0), (1e-9, 0)]:
        m, mean, var = _fit_predict(0, X, y, Xs, noise=noise)
        assert _gate(m) == want, (noise, _gate(m))
        if noise >= 1e-4:
            f = orc.fit(X, y, *THETA, noise)
            mo, vo = orc.predict(X, f, *THETA, Xs)
            np.testing.assert_allclose(mean, mo, rtol=1e-8, atol=1e-9 * np.abs(mo).max())
            np.testing.assert_allclose(var, vo, rtol=1e-8, atol=1e-12)
    m, _, _ = _fit_predict(1, X, y, Xs)
    assert _gate(m) == 0


@pytest.mark.parametrize("mode", [6, 7])
def test_i8_space_time_family(mode):
    """The time factor of the space-time product kernel (scratch.py:506-511) goes through the same generator."""
    rng = np.random.default_rng(3)
    N = 300
    X, y = synthetic.drifter_snapshot(N, config_id=2, seed_offset=4)
    X3 = np.concatenate([rng.uniform(0, 3, (N, 1)), X], 1)
    Xs = synthetic.prediction_grid(X, 20, 21)
    Xs3 = np.concatenate([rng.uniform(0, 3, (Xs.shape[0], 1)), Xs], 1)
    tvar, lt = 1.7, 0.9
    gp.set_predict_i8(mode)
    m = gp.SpaceTimeGP(X3, y, *THETA, tvar, lt, NOISE)
    m.fit()
    mean, var = m.predict(Xs3)
    f = orc.st_fit(X3, y, *THETA, tvar, lt, NOISE)
    mo, vo = orc.st_predict(X3, f, *THETA, tvar, lt, Xs3)
    np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-8, atol=1e-9 * np.abs(mo).max())
    np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-8, atol=1e-12)


def test_i8_long_rows_accumulate_in_segments():
    """Rows longer than 16384 entries are accumulated in segments (int32 range); n = 16 640 has two such row blocks."""
    N = 8320
    X, y = synthetic.drifter_snapshot(N, config_id=5, seed_offset=9)
    Xs = synthetic.prediction_grid(X, 12, 11)
    _, m1, v1 = _fit_predict(1, X, y, Xs)
    for mode in (6, 7):
        m, mm, vv = _fit_predict(mode, X, y, Xs)
        assert _gate(m) == mode
        np.testing.assert_allclose(mm, m1, rtol=1e-9, atol=1e-10 * np.abs(m1).max())
        np.testing.assert_allclose(vv, v1, rtol=2e-9, atol=1e-13)


@pytest.mark.parametrize("mode", [6, 7])
@pytest.mark.parametrize("N,side,theta", [(900, 120, (0.5, 0.8, 0.4)), (2500, 90, THETA), (130, 13, (0.45, 2.2, 0.35))])
def test_i8_zero_slice_skipping_is_bitwise_the_dense_schedule(mode, N, side, theta):
    """Digit slices that are identically zero (covariances of distant points, entries of L^-1 far from the diagonal)
    are neither copied nor multiplied; integer accumulation makes that exact.  Short length scales and grids with
    several column tiles per CTA exercise the sparse schedules: single-stage segments (padded with an empty stage),
    dropped k-steps, both MMA issuers, both panels."""
    import ctypes as C
    from gp2d_b200._lib import lib
    lib.gp2d_dbg_set_i8.restype = C.c_int
    lib.gp2d_dbg_set_i8.argtypes = [C.c_int]
    X, y = synthetic.drifter_snapshot(N, config_id=6, seed_offset=N)
    Xs = gp.as_dev(synthetic.prediction_grid(X, side, side))
    gp.set_predict_i8(mode)
    m = gp.HelmholtzGP(X, y, *theta, NOISE)
    m.fit()
    try:
        lib.gp2d_dbg_set_i8(8)                     # dense schedule: every slice of every k-step
        md, vd = m.predict(Xs)
        lib.gp2d_dbg_set_i8(0)
        for _ in range(3):                         # the sparse schedule is timing dependent in its interleaving only
            ms, vs = m.predict(Xs)
            assert torch.equal(ms, md) and torch.equal(vs, vd)
    finally:
        lib.gp2d_dbg_set_i8(0)


def test_fit_orders_the_observations_in_space():
    """The fit sorts the observations along a Z-order curve (csrc/order.cu): shuffled drifters give the shuffled
    weights and the same prediction as the ordered ones (to rounding: the elimination order is the same), and the
    zero-slice skip finds the same work -- without the sort a shuffled snapshot has no zero slice at all."""
    import ctypes as C
    from gp2d_b200._lib import lib
    lib.gp2d_dbg_i8_counters.restype = C.c_int
    lib.gp2d_dbg_i8_counters.argtypes = [C.POINTER(C.c_ulonglong)]
    cnt = (C.c_ulonglong * 4)()
    N = 1500
    X, y = synthetic.drifter_snapshot(N, config_id=7)
    Xs = gp.as_dev(synthetic.prediction_grid(X, 64, 50))
    p = np.random.default_rng(3).permutation(N)
    Xp, yp = X[p], np.concatenate([y[:N][p], y[N:][p]])
    gp.set_predict_i8(6)
    res = []
    for (Xa, ya) in ((X, y), (Xp, yp)):
        m = gp.HelmholtzGP(Xa, ya, *THETA, NOISE)
        a = m.alpha().cpu().numpy()
        lib.gp2d_dbg_i8_counters(cnt)
        mean, var = m.predict(Xs)
        assert lib.gp2d_dbg_i8_counters(cnt) == 0
        res.append((a, mean.cpu().numpy(), var.cpu().numpy(), int(cnt[0]), int(cnt[2]), m.fit()))
    a0, m0, v0, prod0, dense0, l0 = res[0]
    a1, m1, v1, prod1, dense1, l1 = res[1]
    np.testing.assert_allclose(np.concatenate([a0[:N][p], a0[N:][p]]), a1, rtol=1e-9, atol=1e-11 * np.abs(a0).max())
    np.testing.assert_allclose(m1, m0, rtol=1e-9, atol=1e-11 * np.abs(m0).max())
    np.testing.assert_allclose(v1, v0, rtol=1e-9, atol=1e-13)
    assert abs(l1 - l0) <= 1e-10 * abs(l0)
    assert dense0 == dense1 and abs(prod1 - prod0) <= 0.02 * prod0       # ties of the raster may fall the other way
    assert prod0 < 0.9 * 21 * dense0                                     # and there IS something to skip
    f = orc.fit(Xp, yp, *THETA, NOISE)
    np.testing.assert_allclose(a1, f["alpha"], rtol=1e-8, atol=1e-9 * np.abs(f["alpha"]).max())


def test_spatial_order_is_the_oracles_z_order():
    """The permutation the fit applies (read back through a bring-up export) is oracle.morton_order's, ties included."""
    import ctypes as C
    from gp2d_b200._lib import lib
    lib.gp2d_dbg_fit_order.restype = C.c_int
    lib.gp2d_dbg_fit_order.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(4)
    for N in (255, 256, 1000, 3001):
        X = rng.uniform(-3, 7, (N, 2))
        X[N // 3] = X[N // 2]                          # a tie
        y = rng.normal(size=2 * N)
        m = gp.HelmholtzGP(X, y, *THETA, NOISE)
        m.fit()
        perm = torch.full((N,), -1, dtype=torch.int32, device="cuda")
        rc = lib.gp2d_dbg_fit_order(m.ws.data_ptr(), N, 2, perm.data_ptr(), None)
        torch.cuda.synchronize()
        if N < 256:
            assert rc == 0                              # below the threshold the caller's order is kept
            continue
        assert rc == N
        np.testing.assert_array_equal(perm.cpu().numpy(), orc.morton_order(X))


@pytest.mark.parametrize("case", ["duplicates", "collinear256", "thin257", "tiny"])
def test_spatial_order_degenerate_inputs(case):
    """The Z-order sort of the fit (csrc/order.cu) on inputs with no extent in one or both directions, repeated points,
    and N at the threshold where it switches on: weights in the caller's order, mean and variance against the oracle."""
    rng = np.random.default_rng(1)
    if case == "duplicates":
        X = rng.uniform(0, 5, (300, 2))
        X[50:120] = X[7]
    elif case == "collinear256":
        X = np.stack([np.linspace(0, 9, 256), np.full(256, 3.0)], 1)
    elif case == "thin257":
        X = rng.uniform(0, 5, (257, 2))
        X[:, 1] = X[:, 0] * 1e-12
    else:
        X = rng.uniform(0, 1e-9, (400, 2))
    N = X.shape[0]
    y = rng.normal(size=2 * N) * 0.3
    Xs = rng.uniform(-1, 10, (500, 2))
    theta = (1.3, 2.1, 0.4)
    m = gp.HelmholtzGP(X, y, *theta, NOISE)
    a = m.alpha().cpu().numpy()
    mean, var = m.predict(Xs)
    f = orc.fit(X, y, *theta, NOISE)
    mo, vo = orc.predict(X, f, *theta, Xs)
    np.testing.assert_allclose(a, f["alpha"], rtol=1e-8, atol=1e-9 * np.abs(f["alpha"]).max())
    np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-8, atol=1e-9 * np.abs(mo).max())
    np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-8, atol=1e-12)


def test_i8_concurrent_streams_and_pacing():
    """Three host threads, each fitting and predicting on its own stream at a size where the row-block pacing is on
    (the six slices of Z exceed 64 MB): the kernels queue for the SMs (one CTA takes a whole SM and all of its TMEM),
    CTAs of a launch are not co-resident from the start -- the pacing must give up, not hang -- and every thread gets
    the bits of the sequential run."""
    import threading
    N = 2600
    X, y = synthetic.drifter_snapshot(N, config_id=3)
    Xs = synthetic.prediction_grid(X, 110, 100)
    gp.set_predict_i8(0)
    m0 = gp.HelmholtzGP(X, y, *THETA, NOISE)
    m0.fit()
    assert _gate(m0) == 6
    r0 = m0.predict(Xs)
    torch.cuda.synchronize()
    out = {}

    def work(k):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            m = gp.HelmholtzGP(X, y, *THETA, NOISE)
            m.fit()
            for _ in range(2):
                mean, var = m.predict(Xs)
            s.synchronize()
            out[k] = (mean.clone(), var.clone())

    ths = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    [t.start() for t in ths]
    [t.join(timeout=100) for t in ths]
    assert len(out) == 3
    for k in out:
        assert torch.equal(out[k][0], r0[0]) and torch.equal(out[k][1], r0[1])


def test_i8_host_entry_point_matches_device_path():
    """gp2d_fit_predict_host (numpy in / out) runs the same kernels as fit + predict on device tensors."""
    X, y = synthetic.drifter_snapshot(500, config_id=2, seed_offset=5)
    Xs = synthetic.prediction_grid(X, 25, 24)
    for mode in (0, 1, 7):
        m, mean, var = _fit_predict(mode, X, y, Xs)
        gp.set_predict_i8(mode)
        mu, vh, lml = gp.fit_predict_host(X, y, *THETA, NOISE, Xs)
        np.testing.assert_array_equal(mu, mean)
        np.testing.assert_array_equal(vh, var)
