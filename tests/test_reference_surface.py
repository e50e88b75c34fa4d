"""GPU tests of the reference-facing call surface (myKernel classes, GP_scripts functions,
GP_laser.simLaser, GPRegression-like model, krig workflows) against the golden vectors made
from the reference's own code and against the CPU oracle.  Run on the B200 box."""
import os
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import gp2d_b200 as gp                                        # noqa: E402
from gp2d_b200 import GP_laser, GP_scripts, krig, models, myKernel, synthetic   # noqa: E402
from oracle import gp_oracle as orc                          # noqa: E402


@pytest.fixture(scope="module")
def ks(golden_dir):
    return np.load(os.path.join(golden_dir, "kernel_small.npz"))


# ---- myKernel.py classes (myKernel.py:12-334) --------------------------------------------------
def test_mykernel_class_matches_reference(ks):
    X, X2 = ks["X"], ks["X2"]
    for t, (ldf, lcf, r) in enumerate(ks["thetas"]):
        k = myKernel.myKernel(2, [0, 1], ldf, lcf, r)
        np.testing.assert_allclose(k.K(X), ks["ref_K_class_sym_%d" % t], rtol=0, atol=1e-14)
        np.testing.assert_allclose(k.K(X, X2), ks["ref_K_class_x_%d" % t], rtol=0, atol=1e-14)
        np.testing.assert_allclose(k.Kdiag(X2), ks["ref_Kdiag_%d" % t], rtol=1e-15)
        assert k.Kdiag(X2).shape == ks["ref_Kdiag_%d" % t].shape
        np.testing.assert_allclose(k.param_array, [ldf, lcf, r])
        # the reference's own (incorrect) length-scale integrands, on request
        k.reference_compat = True
        k.update_gradients_full(ks["W_sym"], X, None)
        got = [k.length_df.gradient, k.length_cf.gradient, k.ratio.gradient]
        np.testing.assert_allclose(got, ks["ref_grad_compat_sym_%d" % t], rtol=1e-11, atol=1e-12)
        k.update_gradients_full(ks["W_x"], X, X2)
        got = [k.length_df.gradient, k.length_cf.gradient, k.ratio.gradient]
        np.testing.assert_allclose(got, ks["ref_grad_compat_x_%d" % t], rtol=1e-11, atol=1e-12)
        # default: the analytic derivative
        k.reference_compat = False
        k.update_gradients_full(ks["W_x"], X, X2)
        got = [k.length_df.gradient, k.length_cf.gradient, k.ratio.gradient]
        np.testing.assert_allclose(got, orc.kernel_grad_sums(ks["W_x"], X, X2, ldf, lcf, r), rtol=1e-11, atol=1e-12)


def test_single_component_classes(ks):
    X, X2 = ks["X"], ks["X2"]
    nd = myKernel.nonDivK(2, [0, 1], 1.7)
    np.testing.assert_allclose(nd.K(X, X2), ks["ref_nonDivK"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(nd.Kdiag(X2), ks["ref_nonDivK_diag"], rtol=1e-15)
    nd.reference_compat = True
    nd.update_gradients_full(ks["W_x"], X, X2)
    np.testing.assert_allclose(nd.length.gradient, ks["ref_nonDivK_grad_compat"][0], rtol=1e-11)
    nr = myKernel.nonRotK(2, [0, 1], 0.9)
    np.testing.assert_allclose(nr.K(X, X2), ks["ref_nonRotK"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(nr.Kdiag(X2), ks["ref_nonRotK_diag"], rtol=1e-15)
    nr.reference_compat = True
    nr.update_gradients_full(ks["W_x"], X, X2)
    np.testing.assert_allclose(nr.length.gradient, ks["ref_nonRotK_grad_compat"][0], rtol=1e-11)
    with pytest.raises(AssertionError):
        myKernel.myKernel(3)                                   # input_dim guard (myKernel.py:15)
    with pytest.raises(NotImplementedError):
        nd.gradients_X(ks["W_x"], X, X2)


def test_active_dims_slicing(ks):
    X = ks["X"]
    X3 = np.concatenate([np.full((X.shape[0], 1), 7.0), X], axis=1)       # (t, y, x) rows as in krig
    k = myKernel.myKernel(2, [1, 2], 2.0, 2.0, 0.5)
    np.testing.assert_allclose(k.K(X3), ks["ref_K_class_sym_0"], rtol=0, atol=1e-14)


# ---- GP_scripts.py functions (GP_scripts.py:6-123,44-54) ------------------------------------------
def test_gp_scripts_functions(ks):
    X, X2 = ks["X"], ks["X2"]
    for t, (ldf, lcf, r) in enumerate(ks["thetas"]):
        np.testing.assert_allclose(GP_scripts.myKernel(X, X2, ldf, lcf, r), ks["ref_K_func_x_%d" % t], rtol=0, atol=1e-14)
        K = (r * GP_scripts.compute_K(X[:, 0], X[:, 1], ldf, 1) + (1 - r) * GP_scripts.compute_K(X[:, 0], X[:, 1], lcf, 2))
        np.testing.assert_allclose(K, ks["ref_K_loops_sym_%d" % t], rtol=0, atol=1e-14)
        Ks = (r * GP_scripts.compute_Ks(X[:, 0], X[:, 1], X2[:, 0], X2[:, 1], ldf, 1)
              + (1 - r) * GP_scripts.compute_Ks(X[:, 0], X[:, 1], X2[:, 0], X2[:, 1], lcf, 2))
        np.testing.assert_allclose(Ks, ks["ref_Ks_loops_%d" % t], rtol=0, atol=1e-14)
    blk = GP_scripts.nonDivK(X[0], X[3], 1.7, 1)
    np.testing.assert_allclose(blk, orc.helmholtz_K(X[0:1], X[3:4], 1.7, 1.7, 1.0), rtol=0, atol=1e-15)
    # scalar squared-exponential helpers (GP_scripts.py:67-68,125-142)
    np.testing.assert_allclose(GP_scripts.sqExp(X[:, 0], X[:, 1], X2[:, 0], X2[:, 1], 1.7), ks["ref_sqExp"], rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(GP_scripts.rbf(X[:, 0], X2[:, 0], l=1.3, sigma=0.8, noise=0.05), ks["ref_rbf_x"], rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(GP_scripts.rbf(X[:, 0], X[:, 0], l=1.3, sigma=0.8, noise=0.05), ks["ref_rbf_sym"], rtol=1e-13, atol=1e-16)
    assert GP_scripts.nonDivK(X[0], X[3], 1.7, 0) == pytest.approx(
                                                                   float(np.exp(-np.sum((X[0] - X[3]) ** 2) / (2 * 1.7 ** 2))), rel=1e-13)
    assert GP_scripts.rmse1(np.array([1.0, 3.0]), np.zeros(2)) == pytest.approx(np.sqrt(5.0))


def test_getMean_getCov_inverse_form(golden_dir):
    """The reference's explicit-inverse pipeline (GP_laser.py:177-183) step by step through the
    GP_scripts functions, all on the GPU, against the reference's own result."""
    g = np.load(os.path.join(golden_dir, "simlaser_ts0.npz"))
    X, y, Xs = g["X"], g["y"], g["Xs"][::7]
    l, rate, noise = 2.0, 0.5, float(g["noise"])
    xo, yo = X[:, 0], X[:, 1]
    K = rate * GP_scripts.compute_K(xo, yo, l, 1) + (1 - rate) * GP_scripts.compute_K(xo, yo, l, 2)
    K = K + np.identity(K.shape[0]) * noise
    Ki = gp.spd_inverse(K).cpu().numpy()
    np.testing.assert_allclose(Ki @ K, np.eye(K.shape[0]), atol=1e-9)
    Ks = (rate * GP_scripts.compute_Ks(xo, yo, Xs[:, 0], Xs[:, 1], l, 1)
          + (1 - rate) * GP_scripts.compute_Ks(xo, yo, Xs[:, 0], Xs[:, 1], l, 2))
    f = GP_scripts.getMean(Ks, Ki, y[:, None])
    M = g["Xs"].shape[0]
    ref = np.concatenate([g["ref_mean"][:M][::7], g["ref_mean"][M:][::7]])
    np.testing.assert_allclose(f, ref, rtol=1e-8, atol=1e-8 * np.abs(ref).max())
    # getCov on a single-component kernel: diagonal equals the fused predictive variance
    Xg = Xs[:40]
    ML, Ki1, Ks1 = GP_scripts.getCov(xo[:60], yo[:60], Xg[:, 0], Xg[:, 1], sigma=0.4, divFree=1)
    Kd = orc.helmholtz_K(X[:60], None, 0.4, 0.4, 1.0)
    Ksd = orc.helmholtz_K(Xg, X[:60], 0.4, 0.4, 1.0)
    ref_cov = orc.helmholtz_K(Xg, None, 0.4, 0.4, 1.0) - Ksd @ np.linalg.inv(Kd) @ Ksd.T
    np.testing.assert_allclose(Ks1, Ksd, atol=1e-14)
    np.testing.assert_allclose(ML, ref_cov, rtol=1e-8, atol=1e-9)


# ---- GP_laser.simLaser (GP_laser.py:145-187) ----------------------------------------------------------
@pytest.mark.parametrize("col,ts", [(0, 0), (1, 100)])
def test_simLaser_end_to_end(golden_dir, col, ts):
    t = np.load(os.path.join(golden_dir, "simlaser_tracks.npz"))
    g = np.load(os.path.join(golden_dir, "simlaser_ts%d.npz" % ts))
    tracks = types.SimpleNamespace(lat=t["lat"], lon=t["lon"], u=t["u"], v=t["v"])
    X, Y, uf, vf, xob, yob, u, v, uvar, vvar = GP_laser.simLaser(ts=col, tracks=tracks, return_var=True)
    M = g["Xs"].shape[0]
    assert uf.shape == (51, 51) and X.shape == (51, 51)
    scale = np.abs(g["ref_mean"]).max()
    np.testing.assert_allclose(uf.reshape(-1), g["ref_mean"][:M], rtol=1e-8, atol=1e-8 * scale)
    np.testing.assert_allclose(vf.reshape(-1), g["ref_mean"][M:], rtol=1e-8, atol=1e-8 * scale)
    np.testing.assert_allclose(uvar.reshape(-1), g["ref_var"][:M], rtol=1e-8)
    np.testing.assert_allclose(vvar.reshape(-1), g["ref_var"][M:], rtol=1e-8)
    # the K* weighting exactly as written in GP_laser.py:181
    out = GP_laser.simLaser(ts=col, tracks=tracks, simlaser_compat=True)
    f = np.concatenate([out[2].reshape(-1), out[3].reshape(-1)])
    np.testing.assert_allclose(f, g["ref_mean_simlaser"], rtol=1e-8, atol=1e-8 * scale)


# ---- GPRegression-like model (GP_plots.py:760-770; krig.py:411,438-457,543-544) ----------------------
def _small_problem(N=150, seed=5):
    X, y = synthetic.drifter_snapshot(N, config_id=4, seed_offset=seed)
    return X, y


def test_gpregression_likelihood_gradient_predict():
    X, y = _small_problem()
    k = myKernel.myKernel(2, [0, 1], 1.3, 3.1, 0.2)
    m = models.GPRegression(X, y[:, None], k, noise_var=0.05)
    lo, go = orc.lml_and_grad(X, y, 1.3, 3.1, 0.2, 0.05, jitter=1e-8)
    assert abs(m.log_likelihood() - lo) <= 1e-6 * abs(lo)
    got = [p.gradient for p in m.parameters]
    np.testing.assert_allclose(got, go, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(m.param_array, [1.3, 3.1, 0.2, 0.05])            # GP_plots.py:810-813
    Xs = synthetic.prediction_grid(X, 9, 7)
    mean, var = m.predict(Xs)
    assert mean.shape == (2 * 63, 1) and var.shape == (2 * 63, 1)
    f = orc.fit(X, y, 1.3, 3.1, 0.2, 0.05, jitter=1e-8)
    mo, vo = orc.predict(X, f, 1.3, 3.1, 0.2, Xs, noise=0.05, include_noise=True)   # GPy adds the noise
    np.testing.assert_allclose(mean[:, 0], mo, rtol=1e-8, atol=1e-9 * np.abs(mo).max())
    np.testing.assert_allclose(var[:, 0], vo, rtol=1e-8)


def test_gpregression_optimize_and_pickle(tmp_path):
    X, y = _small_problem(120, seed=9)
    m = models.GPRegression(X, y[:, None], myKernel.myKernel(2, [0, 1], 1.0, 1.0, 0.5))
    ll0 = m.log_likelihood()
    m.optimize_restarts(num_restarts=3, messages=False, verbose=False, seed=4, max_iters=60)
    assert len(m.optimization_runs) == 3
    assert m.log_likelihood() > ll0
    best = min(r.f_opt for r in m.optimization_runs)
    assert m.objective_function() == pytest.approx(best, rel=1e-9, abs=1e-9)
    ldf, lcf, r, nz = m.param_array
    assert ldf > 0 and lcf > 0 and 0 < r < 1 and nz > 0
    # at the optimum the gradient in the unconstrained space is small
    lo, go = orc.lml_and_grad(X, y, ldf, lcf, r, nz, jitter=1e-8)
    assert abs(m.log_likelihood() - lo) <= 1e-6 * max(abs(lo), 1.0)
    p = str(tmp_path / "model.pkl")
    m.pickle(p)
    m2 = models.load(p)
    np.testing.assert_array_equal(m2.param_array, m.param_array)
    assert m2.log_likelihood() == pytest.approx(m.log_likelihood(), rel=1e-12)
    assert len(m2.optimization_runs) == 3
    Xs = synthetic.prediction_grid(X, 5, 5)
    a, b = m.predict(Xs), m2.predict(Xs)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_restart_sharding_is_a_partition():
    """Restart r runs on rank r % world with a seed that depends only on r, so the union over
    ranks equals the single-rank run (SURVEY.md §8e axis 2)."""
    X, y = _small_problem(80, seed=2)
    def run(rank, world):
        m = models.GPRegression(X, y[:, None], myKernel.myKernel(2, [0, 1], 1.0, 1.0, 0.5))
        m.optimize_restarts(num_restarts=4, verbose=False, seed=11, max_iters=25, rank=rank, world=world)
        return [r.f_opt for r in m.optimization_runs]
    full = run(0, 1)
    a, b = run(0, 2), run(1, 2)
    assert len(a) == 2 and len(b) == 2
    np.testing.assert_allclose(sorted(a + b), sorted(full), rtol=1e-9, atol=1e-9)
    # concurrent restarts on separate streams / workspaces: the same runs, in restart order
    m = models.GPRegression(X, y[:, None], myKernel.myKernel(2, [0, 1], 1.0, 1.0, 0.5))
    m.optimize_restarts(num_restarts=4, verbose=False, seed=11, max_iters=25, parallel=3)
    np.testing.assert_allclose([r.f_opt for r in m.optimization_runs], full, rtol=1e-9, atol=1e-9)
    assert m.objective_function() == pytest.approx(min(full), rel=1e-9, abs=1e-9)


# ---- krig workflows (krig.py:259-418,430-468,471-574,578-645) ----------------------------------------------
def test_krig_workflow(tmp_path):
    rng = np.random.default_rng(3)
    nt, nd = 4, 60
    lat = 28.8 + rng.uniform(0.0, 0.12, size=(1, nd)) + np.zeros((nt, 1))
    lon = -88.6 + rng.uniform(0.0, 0.12, size=(1, nd)) + np.zeros((nt, 1))
    lat[2, 5] = np.nan
    lon[2, 5] = np.nan
    time = np.arange(nt) * 0.25
    u = 0.2 * np.sin(30 * (lat - 28.8)) + rng.normal(0, 0.02, size=lat.shape)
    v = 0.2 * np.cos(30 * (lon + 88.6)) + rng.normal(0, 0.02, size=lat.shape)
    out = str(tmp_path / "rbfModel")
    model = krig.kriging(0, nt, sample_step=-1, skip=2, output=out, kernelType=4,
                         data=(time, lat, lon, v, u, np.full(nd, float(nt))))
    assert os.path.isfile(out + "_combined.pkl") and os.path.isfile(out + ".mat")
    import scipy.io as sio
    mat = sio.loadmat(out + ".mat")
    n_obs = mat["Xo"].shape[0]
    assert mat["obs"].shape == (2 * n_obs, 1) and mat["Xo"].shape[1] == 3
    assert n_obs == nt * 30                      # drifters 0,2,..,58; the NaN sits on a test drifter
    ll0 = model.log_likelihood()
    m2 = krig.runRestarts(out, nres=2, seed=1, max_iters=30)
    assert m2.log_likelihood() >= ll0
    Xp, V, U, VVar, UVar = krig.predict(out, tlim=[0, 0.5], ylim=[0, 12], xlim=[0, 12], dt=0.25, dx=1.0)
    assert V.shape == U.shape == VVar.shape == UVar.shape and V.ndim == 3
    assert np.all(VVar > 0)
    from scipy.io import netcdf_file
    f = netcdf_file(out + ".nc", "r", mmap=False)
    np.testing.assert_allclose(f.variables["u"].data, U.astype(np.float32), rtol=1e-6)
    np.testing.assert_allclose(f.variables["hyperparam_v"].data, m2.param_array.astype(np.float32), rtol=1e-6)
    f.close()
    res = krig.predictTest(out)
    assert res["Vp"].shape[0] == mat["Xt"].shape[0]
    rv, ru = krig.getRMSE(out)
    assert 0 < rv < 0.5 and 0 < ru < 0.5
    with pytest.raises(ValueError):
        krig.make_kernel(5)


# ---- physical invariants and the reference's own end-to-end example ---------------------------------------------
def test_divergence_free_kernel_gives_divergence_free_mean():
    """SURVEY.md §4 'physical invariants' (report.tex:74-82): the posterior mean of the
    divergence-free kernel has zero divergence, that of the curl-free kernel zero curl --
    whatever the data.  Checked with centred differences on a fine grid against the size of the
    velocity gradients themselves."""
    rng = np.random.default_rng(7)
    X = rng.uniform(-1, 1, (60, 2))
    y = rng.normal(size=120)                                    # arbitrary data: the property is structural
    h = 1e-3
    g = np.linspace(-0.8, 0.8, 21)
    GX, GY = np.meshgrid(g, g)
    P = np.stack([GX.ravel(), GY.ravel()], axis=1)
    pts = np.concatenate([P + [h, 0], P - [h, 0], P + [0, h], P - [0, h]])
    for ratio, which in ((1.0, "div"), (0.0, "curl")):
        m = gp.HelmholtzGP(X, y, 0.5, 0.5, ratio, 0.01)
        m.fit()
        mean, _ = m.predict(pts)
        mean = mean.cpu().numpy()
        n = pts.shape[0]
        u, v = mean[:n].reshape(4, -1), mean[n:].reshape(4, -1)
        ux, uy = (u[0] - u[1]) / (2 * h), (u[2] - u[3]) / (2 * h)
        vx, vy = (v[0] - v[1]) / (2 * h), (v[2] - v[3]) / (2 * h)
        scale = np.sqrt(np.mean(ux ** 2 + uy ** 2 + vx ** 2 + vy ** 2))
        resid = (ux + vy) if which == "div" else (vx - uy)
        assert np.max(np.abs(resid)) < 1e-5 * scale, (which, np.max(np.abs(resid)), scale)
        other = (vx - uy) if which == "div" else (ux + vy)
        assert np.max(np.abs(other)) > 1e-2 * scale                 # the other quantity is not constrained


def test_run_example3_reconstructs_the_synthetic_field():
    """GP_plots.run_example3 (GP_plots.py:705-779) end to end: product-of-RBF models per component and
    the joint Helmholtz model after optimize_restarts; the vector kernel, which knows the field is
    divergence-free, reconstructs it better from the same samples."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "run_example3", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "run_example3.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    rmsug, rmsvg, rmsfu, rmsfv, u_HP, v_HP, f_HP = ex.run_example3(nsamples=30, divFree=1, rms=1, num_restarts=3, seed=1)
    x, y, phi, xm, ym, um, vm = GP_scripts.generate_2D_gaussian(1)
    amp = np.sqrt(np.mean(um ** 2 + vm ** 2))
    assert rmsfu < 0.05 * amp and rmsfv < 0.05 * amp
    assert rmsfu + rmsfv < rmsug + rmsvg
    assert u_HP.size == 5 and f_HP.size == 4
    # the kernel has no amplitude parameter: the prior variances of its two parts are ratio / l_df^2 and
    # (1 - ratio) / l_cf^2, and the fit puts (almost) all of it on the divergence-free part
    w_df, w_cf = f_HP[2] / f_HP[0] ** 2, (1 - f_HP[2]) / f_HP[1] ** 2
    assert w_df > 100 * w_cf


def test_product_of_rbf_kernels_is_an_ard_rbf(tmp_path):
    from gp2d_b200 import kern
    rng = np.random.default_rng(2)
    X = rng.uniform(0, 1, (50, 2))
    y = np.sin(3 * X[:, 0]) * np.cos(2 * X[:, 1]) + 0.01 * rng.normal(size=50)
    k = kern.RBF(1, active_dims=[0], variance=2.0, lengthscale=0.3) * kern.RBF(1, active_dims=[1], variance=1.5, lengthscale=0.5)
    np.testing.assert_allclose(k.K(X), orc.rbf_sum_K(X, None, [3.0], [[0.3, 0.5]]), rtol=1e-13, atol=1e-15)
    m = models.GPRegression(X, y[:, None], k, noise_var=0.01)
    lo, go = orc.rbf_lml_and_grad(X, y, [3.0], [[0.3, 0.5]], 0.01, jitter=1e-8)
    assert abs(m.log_likelihood() - lo) <= 1e-6 * abs(lo)
    # d/dv1 = v2 d/dvar, d/dv2 = v1 d/dvar (the product's variance is v1 v2)
    np.testing.assert_allclose([p.gradient for p in m.parameters], [go[0] * 1.5, go[1], go[0] * 2.0, go[2], go[3]], rtol=1e-6, atol=1e-7)
    p = str(tmp_path / "prod.pkl")
    m.pickle(p)
    m2 = models.load(p)
    np.testing.assert_array_equal(m2.param_array, m.param_array)
    assert m2.log_likelihood() == pytest.approx(m.log_likelihood(), rel=1e-12)


def test_laser_workflow_on_synthetic_tracks():
    """GP_laser.laser (GP_laser.py:16-142) with an injected tracks object: the reference's explicit-
    inverse pipeline (compute_K, inv, compute_Ks, getMean, diag of Kss - Ks Ki Ks^T) restated by the
    oracle on the same filtered observations."""
    rng = np.random.default_rng(12)
    nd, nt = 45, 40
    lat = 28.8 + rng.uniform(0.0, 0.25, size=(nd, 1)) + 0.0005 * np.arange(nt)[None, :]
    lon = -88.55 + rng.uniform(0.0, 0.25, size=(nd, 1)) + 0.0004 * np.arange(nt)[None, :]
    u = 0.3 * np.sin(25 * (lat - 28.8)) + rng.normal(0, 0.02, size=lat.shape)
    v = 0.3 * np.cos(25 * (lon + 88.55)) + rng.normal(0, 0.02, size=lat.shape)
    drogue = np.ones((nd, nt))
    drogue[3, :] = 0                      # an undrogued drifter is dropped
    u[5, 22] = 3.5                        # a spike above 2 m/s is dropped
    lat[7, 21], lon[7, 21] = np.nan, np.nan
    tr = types.SimpleNamespace(lat=lat, lon=lon, u=u, v=v, drogueStat=drogue, time=np.arange(nt) * 900.0)
    out = GP_laser.laser(ts=20, nsteps=4, l_df=5, l_cf=4, rate=0.4, noise=0.0025, nsamples=1, tracks=tr)
    x, y, uf, vf, xo, yo, uo, vo, uvar, vvar, xt, yt, ut, vt, uft, vft = out
    n_all = (nd - 1) * 4 - 2              # one drifter out, one spike, one missing position
    assert xo.size == len(range(0, n_all, 3)) and xt.size == n_all - xo.size
    assert xo.min() >= 2.0 and uf.shape == (y.size, x.size) and uvar.shape == uf.shape
    X = np.stack([xo, yo], axis=1)
    Xg, Yg = np.meshgrid(x, y)
    Xs = np.stack([Xg.reshape(-1), Yg.reshape(-1)], axis=1)
    mo, vo_ = orc.fit_predict_inverse_form(X, np.concatenate([uo, vo]), 5, 4, 0.4, 0.0025, Xs)
    M = Xs.shape[0]
    scale = np.abs(mo).max()
    np.testing.assert_allclose(uf.reshape(-1), mo[:M], rtol=1e-7, atol=1e-8 * scale)
    np.testing.assert_allclose(vf.reshape(-1), mo[M:], rtol=1e-7, atol=1e-8 * scale)
    np.testing.assert_allclose(uvar.reshape(-1), vo_[:M], rtol=1e-6, atol=1e-10)
    mt, _ = orc.fit_predict_inverse_form(X, np.concatenate([uo, vo]), 5, 4, 0.4, 0.0025, np.stack([xt, yt], axis=1), want_var=False)
    np.testing.assert_allclose(np.concatenate([uft, vft]), mt, rtol=1e-7, atol=1e-8 * scale)
    assert GP_laser.laser2 is GP_laser.laser


def test_model_retries_with_jitter_like_gpys_jitchol():
    """Duplicate observation sites and no noise: K + 1e-8 I is numerically singular.  GPy's jitchol
    (GPy/util/linalg.py) then adds 1e-6 mean(diag K) 10^k to the diagonal and warns; so does the model
    here, instead of reporting -inf."""
    rng = np.random.default_rng(5)
    X = rng.uniform(0, 5, (40, 2))
    X = np.concatenate([X, X, X])                                  # every site three times
    y = rng.normal(size=2 * X.shape[0])
    with pytest.warns(RuntimeWarning, match="Added jitter"):
        m = models.GPRegression(X, y[:, None], myKernel.myKernel(2, [0, 1], 8.0, 8.0, 0.5), noise_var=1e-300, jitter=0.0)
    assert np.isfinite(m.log_likelihood())
    mean, var = m.predict(X[:5], refined=False)
    assert np.all(np.isfinite(mean)) and np.all(var >= 0)
