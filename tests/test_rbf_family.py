"""GPU tests of the scalar ARD-RBF sum family (krig.scikit_prior,
kernelType=1) against golden vectors from LIVE scikit-learn (tests/golden/make_golden_sklearn.py)
and against the CPU oracle."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import gp2d_b200 as gp                                        # noqa: E402
from gp2d_b200 import kern, krig, models                     # noqa: E402
from gp2d_b200.sklearn_like import GaussianProcessRegressor, kernels   # noqa: E402
from oracle import gp_oracle as orc                          # noqa: E402


def _hp_split(HP):
    Q = (HP.size - 1) // 4
    return [HP[4 * q] for q in range(Q)], [HP[4 * q + 1:4 * q + 4] for q in range(Q)], float(HP[-1])


@pytest.fixture(params=["sklearn_rbf1", "sklearn_rbf2"])
def gold(request, golden_dir):
    return np.load(os.path.join(golden_dir, request.param + ".npz"))


def test_kernel_build_matches_sklearn(gold):
    var, ls, noise = _hp_split(gold["HP"])
    K = gp.rbf_K(gold["XT"][:40], gold["Xg"][:30], var, ls).cpu().numpy()
    np.testing.assert_allclose(K, gold["ref_K_signal"], rtol=1e-12, atol=1e-16)
    Kt = gp.rbf_K(gold["XT"][:50], None, var, ls, diag_add=noise).cpu().numpy()
    np.testing.assert_allclose(Kt, gold["ref_K_train"], rtol=1e-12, atol=1e-16)


@pytest.mark.parametrize("N,M,D,Q", [(1, 1, 1, 1), (5, 3, 2, 1), (130, 257, 3, 2), (300, 129, 4, 4), (129, 1000, 3, 3)])
def test_kernel_build_ragged_vs_oracle(N, M, D, Q):
    rng = np.random.default_rng(N + M)
    X, X2 = rng.uniform(0, 10, (N, D)), rng.uniform(0, 10, (M, D))
    var, ls = rng.uniform(0.1, 2, Q), rng.uniform(0.5, 6, (Q, D))
    np.testing.assert_allclose(gp.rbf_K(X, X2, var, ls).cpu().numpy(), orc.rbf_sum_K(X, X2, var, ls), rtol=1e-12, atol=1e-15)
    W = rng.normal(size=(N, M))
    g = gp.rbf_grad_sums(W, X, X2, var, ls).cpu().numpy()
    np.testing.assert_allclose(g, orc.rbf_kernel_grad_sums(W, X, X2, var, ls), rtol=1e-10, atol=1e-12)


def test_sklearn_pipeline_as_scikit_prior_builds_it(gold):
    """krig.py:174-194 with the look-alike classes: tolerances of BASELINE.json (1e-8 relative on
    mean and variance, 1e-6 on the log-likelihood)."""
    HP = gold["HP"]
    k = HP[0] * kernels.RBF(length_scale=[HP[1], HP[2], HP[3]])
    if HP.size - 1 > 5:
        k = k + HP[4] * kernels.RBF(length_scale=[HP[5], HP[6], HP[7]])
    k = k + kernels.WhiteKernel(noise_level=HP[-1])
    m = GaussianProcessRegressor(kernel=k, optimizer=None).fit(gold["XT"], gold["u"][:, None])
    U, Ustd = m.predict(gold["Xg"], return_std=True)
    assert U.shape == (gold["Xg"].shape[0], 1) and Ustd.shape == (gold["Xg"].shape[0],)
    scale = np.abs(gold["ref_mean"]).max()
    np.testing.assert_allclose(U[:, 0], gold["ref_mean"], rtol=1e-8, atol=1e-8 * scale)
    np.testing.assert_allclose(Ustd ** 2, gold["ref_var"], rtol=1e-8)
    assert abs(m.log_marginal_likelihood_value_ - float(gold["ref_lml"])) <= 1e-6 * abs(float(gold["ref_lml"]))
    with pytest.raises(NotImplementedError):
        GaussianProcessRegressor(kernel=k, optimizer="my_optimizer")


def test_lml_gradient_matches_sklearn(gold):
    var, ls, noise = _hp_split(gold["HP"])
    g = gp.ScalarGP(gold["XT"], gold["u"], var, ls, noise, jitter=float(gold["sklearn_alpha"]))
    lml, grad = g.lml_and_grad()
    assert abs(lml - float(gold["ref_lml"])) <= 1e-6 * abs(float(gold["ref_lml"]))
    theta = np.exp(gold["ref_theta"])                     # sklearn differentiates w.r.t. log(theta)
    np.testing.assert_allclose(grad * theta, gold["ref_grad_logtheta"], rtol=1e-6, atol=1e-7)
    # a valid fit state is left behind
    mean, _ = g.predict(gold["Xg"])
    np.testing.assert_allclose(mean.cpu().numpy(), gold["ref_mean"], rtol=1e-8, atol=1e-8 * np.abs(gold["ref_mean"]).max())


@pytest.mark.parametrize("N,M,D,Q", [(1, 2, 3, 1), (127, 1, 3, 2), (128, 128, 2, 1), (700, 3001, 3, 2), (1300, 517, 4, 3)])
def test_fit_predict_vs_oracle(N, M, D, Q):
    rng = np.random.default_rng(100 + N)
    X = rng.uniform(0, 12, (N, D))
    y = np.sin(X[:, 0] / 2.0) + 0.1 * rng.normal(size=N)
    Xs = rng.uniform(0, 12, (M, D))
    var, ls, noise = rng.uniform(0.2, 1.5, Q), rng.uniform(1.0, 5.0, (Q, D)), 0.01
    g = gp.ScalarGP(X, y, var, ls, noise, jitter=1e-8)
    lml = g.fit()
    mean, v = g.predict(Xs, include_noise=True)
    f = orc.rbf_fit(X, y, var, ls, noise, jitter=1e-8)
    mo, vo = orc.rbf_predict(X, f, var, ls, Xs, var_add=noise)
    assert abs(lml - f["lml"]) <= 1e-6 * max(abs(f["lml"]), 1.0)
    np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-8, atol=1e-8 * max(np.abs(mo).max(), 1e-3))
    np.testing.assert_allclose(v.cpu().numpy(), vo, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(g.alpha().cpu().numpy(), f["alpha"], rtol=1e-7, atol=1e-8 * np.abs(f["alpha"]).max())
    # grid partition / split invariance, bit for bit
    if M > 200:
        m1, v1 = g.predict(Xs[:200], include_noise=True)
        assert torch.equal(m1, mean[:200]) and torch.equal(v1, v[:200])
    # shipping the predict state to another workspace
    h = gp.ScalarGP(X * 0 + 1.0, y * 0, var, ls, noise, jitter=1e-8)
    h.predict_state().copy_(g.predict_state())
    h.fitted = True
    m2, v2 = h.predict(Xs, include_noise=True)
    assert torch.equal(m2, mean) and torch.equal(v2, v)


def test_gpy_style_rbf_model_and_restarts(tmp_path):
    """GPy.kern.RBF(input_dim=3, ARD=True) summed twice -> GPRegression -> optimize_restarts ->
    predict, as krig.kriging / runRestarts / predict do for kernelType=1 (krig.py:388-412,450,543)."""
    rng = np.random.default_rng(5)
    N = 160
    X = np.stack([rng.uniform(0, 6, N), rng.uniform(0, 10, N), rng.uniform(0, 10, N)], axis=1)
    y = (0.3 * np.sin(X[:, 1] / 2.0) * np.cos(X[:, 2] / 3.0) + 0.02 * rng.normal(size=N))[:, None]
    k2 = kern.RBF(input_dim=3, ARD=True)
    k = k2.copy() + k2
    m = models.GPRegression(X, y, k)
    assert m.param_array.size == 9 and m.parameter_names()[-1] == "Gaussian_noise.variance"
    var, ls = k.rbf_params()
    lo, go = orc.rbf_lml_and_grad(X, y, [1.0, 1.0], np.ones((2, 3)), 1.0, jitter=1e-8)
    assert abs(m.log_likelihood() - lo) <= 1e-6 * abs(lo)
    np.testing.assert_allclose([p.gradient for p in m.parameters], go, rtol=1e-6, atol=1e-7)
    ll0 = m.log_likelihood()
    m.optimize_restarts(num_restarts=2, verbose=False, seed=3, max_iters=40)
    assert m.log_likelihood() > ll0
    p = str(tmp_path / "m_v.pkl")
    m.pickle(p)
    m2 = models.load(p)
    np.testing.assert_array_equal(m2.param_array, m.param_array)
    Xs = np.stack([np.full(50, 3.0), rng.uniform(0, 10, 50), rng.uniform(0, 10, 50)], axis=1)
    a, b = m.predict(Xs), m2.predict(Xs)
    np.testing.assert_array_equal(a[0], b[0])
    HP = m.param_array
    f = orc.rbf_fit(X, y, [HP[0], HP[4]], [HP[1:4], HP[5:8]], HP[8], jitter=1e-8)
    mo, vo = orc.rbf_predict(X, f, [HP[0], HP[4]], [HP[1:4], HP[5:8]], Xs, var_add=HP[8])
    np.testing.assert_allclose(a[0][:, 0], mo, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(a[1][:, 0], vo, rtol=1e-7)
    # non-ARD kernel: one shared length scale, gradient summed over the dimensions
    m3 = models.GPRegression(X, y, kern.RBF(3, variance=0.5, lengthscale=2.0), noise_var=0.1)
    lo3, go3 = orc.rbf_lml_and_grad(X, y, [0.5], [[2.0, 2.0, 2.0]], 0.1, jitter=1e-8)
    np.testing.assert_allclose([p.gradient for p in m3.parameters], [go3[0], go3[1:4].sum(), go3[4]], rtol=1e-6, atol=1e-7)


def test_krig_scalar_workflow_and_scikit_prior(tmp_path):
    """kriging(kernelType=1) -> runRestarts per component -> predict -> scikit_prior, the
    reference's production chain (runKrig.py:36; krig.py:430-468,471-574; runPredict.py:48)."""
    rng = np.random.default_rng(8)
    nt, nd = 5, 40
    lat = 28.8 + rng.uniform(0.0, 0.12, size=(1, nd)) + np.zeros((nt, 1))
    lon = -88.6 + rng.uniform(0.0, 0.12, size=(1, nd)) + np.zeros((nt, 1))
    time = np.arange(nt) * 0.25
    u = 0.2 * np.sin(30 * (lat - 28.8)) + rng.normal(0, 0.02, size=lat.shape)
    v = 0.2 * np.cos(30 * (lon + 88.6)) + rng.normal(0, 0.02, size=lat.shape)
    d = tmp_path / "run"
    d.mkdir()
    out = str(d / "rbfModel")
    mv, mu = krig.kriging(0, nt, sample_step=-1, skip=2, nKernels=2, output=out, kernelType=1,
                          data=(time, lat, lon, v, u, np.full(nd, float(nt))))
    assert os.path.isfile(out + "_v.pkl") and os.path.isfile(out + "_u.pkl")
    assert mv.param_array.size == 9
    import scipy.io as sio
    mat = sio.loadmat(out + ".mat")
    assert mat["obs"].shape == (mat["Xo"].shape[0], 2)                 # [v, u] columns (krig.py:378)
    for comp in ("_v", "_u"):
        krig.runRestarts(out + comp, nres=1, seed=2, max_iters=30)
    Xp, V, U, VVar, UVar = krig.predict(out, tlim=[0, 1.0], ylim=[0, 12], xlim=[0, 12], dt=0.5, dx=2.0)
    assert V.shape == U.shape == VVar.shape and np.all(UVar > 0)
    res = krig.predictTest(out)
    assert res["Vp"].shape == (mat["Xt"].shape[0], 1) and np.all(res["UpVar"] > 0)
    rv, ru = krig.getRMSE(out)
    assert 0 < rv < 0.5 and 0 < ru < 0.5
    # scikit_prior on an explicit window: compare with the oracle on the same windowed data
    outFile, Uh, Uvar = krig.scikit_prior(out, varname='u', dt=0.5, tlim=6, xlim=[0, 12], ylim=[0, 12], dx=2.0)
    assert os.path.isfile(outFile)
    HP = models.load(out + "_u.pkl").param_array
    XT = np.concatenate([mat["Xo"], mat["Xt"]], axis=0)
    keep = (XT[:, 2] >= 0 - 3) & (XT[:, 2] <= 12 + 3) & (np.abs(XT[:, 0] - 0.5) <= 6)
    uu = np.concatenate([mat["obs"][:, 1], mat["test_points"][:, 1]])
    Xg, tc, yg, xg = krig.getGrid([0.5, 1.5], [0, 12], [0, 12], 1, 2.0)
    f = orc.rbf_fit(XT[keep], uu[keep], [HP[0], HP[4]], [HP[1:4], HP[5:8]], HP[8], jitter=1e-10)
    mo, vo = orc.rbf_predict(XT[keep], f, [HP[0], HP[4]], [HP[1:4], HP[5:8]], Xg, var_add=HP[8])
    np.testing.assert_allclose(Uh.reshape(-1), mo, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(Uvar.reshape(-1), vo, rtol=1e-7)
    from scipy.io import netcdf_file
    fnc = netcdf_file(outFile, "r", mmap=False)
    np.testing.assert_allclose(fnc.variables["u"].data.reshape(-1), mo.astype(np.float32), rtol=1e-5, atol=1e-6)
    fnc.close()


def test_one_dimensional_track_model_with_gpy_idioms():
    """The per-drifter 1-D GP of laser_io_methods.interp_kriging (laser_io_methods.py:464-531):
    RBF(input_dim=1, variance, lengthscale) over time, noise set by assignment, predict, optimize,
    then ``model.rbf.lengthscale[0]`` / ``model.Gaussian_noise[0]`` read back."""
    rng = np.random.default_rng(3)
    t = np.sort(rng.uniform(0, 48, 90))[:, None]
    lon = (-88.0 + 0.01 * t[:, 0] + 0.02 * np.sin(t[:, 0] / 6.0) + 1e-4 * rng.normal(size=90))[:, None]
    k = kern.RBF(input_dim=1, variance=1159.68, lengthscale=4.5)
    m = models.GPRegression(t, lon, k)
    m.Gaussian_noise = 1.75598244486e-07
    assert m.Gaussian_noise[0] == 1.75598244486e-07
    tg = np.arange(1.0, 47.0, 0.25)[:, None]
    mean, var = m.predict(tg)
    f = orc.rbf_fit(t, lon[:, 0], [1159.68], [[4.5]], 1.75598244486e-07, jitter=1e-8)
    mo, vo = orc.rbf_predict(t, f, [1159.68], [[4.5]], tg, var_add=1.75598244486e-07)
    # cond(K) ~ 1e10 here (variance 1e3 against noise 1e-7): compare at the accuracy the problem allows
    np.testing.assert_allclose(mean[:, 0], mo, rtol=0, atol=1e-4)
    assert np.all(var > 0)
    m.optimize(max_iters=20)
    assert m.rbf.lengthscale[0] > 0 and m.rbf.variance[0] > 0 and m.Gaussian_noise[0] > 0
    assert m.rbf is m.kern


def test_sklearn_default_optimizer_reaches_sklearns_optimum(golden_dir):
    """scikit-learn's default (optimizer='fmin_l_bfgs_b'; testKrig.py:139-140): maximise the LML over
    log(theta) from the kernel's initial values.  Same start, same bounds, same optimiser: the search
    ends in scikit-learn's optimum (golden from live scikit-learn)."""
    g = np.load(os.path.join(golden_dir, "sklearn_rbf_opt.npz"))
    s = g["start"]
    k = s[0] * kernels.RBF(length_scale=[s[1], s[2], s[3]]) + kernels.WhiteKernel(noise_level=s[4])
    m = GaussianProcessRegressor(kernel=k, n_restarts_optimizer=0).fit(g["XT"], g["u"])
    assert abs(m.log_marginal_likelihood_value_ - float(g["ref_lml"])) <= 1e-6 * abs(float(g["ref_lml"]))
    np.testing.assert_allclose(m.theta_, g["ref_theta"], rtol=0, atol=2e-3)
    U, Ustd = m.predict(g["Xg"], return_std=True)
    assert U.shape == (g["Xg"].shape[0],)
    np.testing.assert_allclose(U, g["ref_mean"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(Ustd ** 2, g["ref_var"], rtol=1e-3)
    # the initial kernel object is left untouched, as in scikit-learn
    assert k.k2.noise_level == s[4]
    # restarts only ever improve on it
    m2 = GaussianProcessRegressor(kernel=k, n_restarts_optimizer=2, random_state=0).fit(g["XT"], g["u"])
    assert m2.log_marginal_likelihood_value_ >= m.log_marginal_likelihood_value_ - 1e-9
