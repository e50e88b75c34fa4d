"""GPU tests of the space-time product kernel Kt(t) * Helmholtz(y, x) (SURVEY.md §8f rank 1;
scratch.py:506-508, myKernel.py:337-363) against the CPU oracle.  The reference path is dead code
upstream, so the oracle restates the formulas ("parity unpinned"); it is checked here against
central differences and against the pinned Helmholtz / RBF pieces it is built from."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import gp2d_b200 as gp                                        # noqa: E402
from gp2d_b200 import models, myKernel                       # noqa: E402
from oracle import gp_oracle as orc                          # noqa: E402

THETAS = [(0.9, 0.9, 1.0, 5.0, 1.0), (1.3, 3.1, 0.2, 0.7, 2.5), (2.0, 1.1, 0.0, 1.0, 0.4)]


def _data(N, M, seed):
    rng = np.random.default_rng(seed)
    X3 = np.stack([rng.uniform(0, 6, N), rng.uniform(0, 10, N), rng.uniform(0, 10, N)], axis=1)
    Xs3 = np.stack([rng.uniform(0, 6, M), rng.uniform(0, 10, M), rng.uniform(0, 10, M)], axis=1)
    y = np.concatenate([np.sin(X3[:, 1] / 2) * np.cos(X3[:, 0] / 3), np.cos(X3[:, 2] / 2)]) + 0.05 * rng.normal(size=2 * N)
    return X3, Xs3, y


@pytest.mark.parametrize("theta", THETAS)
@pytest.mark.parametrize("N,M", [(1, 1), (9, 6), (130, 257)])
def test_kernel_and_gradient_sums(theta, N, M):
    X3, Xs3, _ = _data(N, M, N + M)
    np.testing.assert_allclose(gp.st_K(X3, Xs3, *theta).cpu().numpy(), orc.st_K(X3, Xs3, *theta), rtol=0, atol=1e-14 * theta[3] * 4)
    np.testing.assert_allclose(gp.st_K(X3, None, *theta, diag_add=0.3).cpu().numpy(),
                               orc.st_K(X3, None, *theta) + 0.3 * np.eye(2 * N), rtol=0, atol=1e-14 * theta[3] * 4)
    W = np.random.default_rng(3).normal(size=(2 * N, 2 * M))
    g = gp.st_grad_sums(W, X3, Xs3, *theta).cpu().numpy()
    go = np.array([np.sum(d * W) for d in orc.st_dK(X3, Xs3, *theta)])
    np.testing.assert_allclose(g, go, rtol=1e-10, atol=1e-11)


def test_oracle_gradient_is_the_derivative():
    X3, Xs3, _ = _data(12, 7, 5)
    W = np.random.default_rng(4).normal(size=(24, 14))
    theta = np.array([1.3, 3.1, 0.2, 0.7, 2.5])
    go = np.array([np.sum(d * W) for d in orc.st_dK(X3, Xs3, *theta)])
    f = lambda th: np.sum(orc.st_K(X3, Xs3, *th) * W)
    fd = np.array([(f(theta + 1e-6 * e) - f(theta - 1e-6 * e)) / 2e-6 for e in np.eye(5)])
    np.testing.assert_allclose(go, fd, rtol=1e-7, atol=1e-8)
    # with the time factor switched off (equal times, unit variance) it is the pinned Helmholtz kernel
    X0 = X3.copy(); X0[:, 0] = 1.0
    np.testing.assert_array_equal(orc.st_K(X0, None, 1.3, 3.1, 0.2, 1.0, 9.0), orc.helmholtz_K(X0[:, 1:], None, 1.3, 3.1, 0.2))


@pytest.mark.parametrize("N,M,theta", [(1, 3, THETAS[0]), (65, 130, THETAS[1]), (300, 517, THETAS[2]), (700, 2601, THETAS[1])])
def test_fit_predict_lml_grad_vs_oracle(N, M, theta):
    X3, Xs3, y = _data(N, M, 100 + N)
    noise = 0.02
    g = gp.SpaceTimeGP(X3, y, *theta, noise, jitter=1e-8)
    lml, grad = g.lml_and_grad()
    mean, var = g.predict(Xs3, include_noise=True)
    f = orc.st_fit(X3, y, *theta, noise, jitter=1e-8)
    mo, vo = orc.st_predict(X3, f, *theta, Xs3, var_add=noise)
    lo, go = orc.st_lml_and_grad(X3, y, *theta, noise, jitter=1e-8)
    assert abs(lml - lo) <= 1e-6 * max(abs(lo), 1.0)
    np.testing.assert_allclose(grad, go, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-8, atol=1e-9 * max(np.abs(mo).max(), 1e-3))
    np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(g.alpha().cpu().numpy(), f["alpha"], rtol=1e-7, atol=1e-8 * np.abs(f["alpha"]).max())
    if M > 300:
        m1, v1 = g.predict(Xs3[:300], include_noise=True)
        assert torch.equal(m1, torch.cat([mean[:300], mean[M:M + 300]])) and torch.equal(v1, torch.cat([var[:300], var[M:M + 300]]))


def test_time_factor_off_matches_the_helmholtz_path():
    """Equal times and tvar = 1: the space-time entry points reproduce gp2d_fit / gp2d_predict bit for bit."""
    X3, Xs3, y = _data(200, 333, 9)
    X3[:, 0] = 2.0
    Xs3[:, 0] = 2.0
    a = gp.SpaceTimeGP(X3, y, 1.3, 3.1, 0.2, 1.0, 4.0, 0.05)
    b = gp.HelmholtzGP(X3[:, 1:], y, 1.3, 3.1, 0.2, 0.05)
    assert a.fit() == b.fit()
    ma, va = a.predict(Xs3)
    mb, vb = b.predict(Xs3[:, 1:])
    assert torch.equal(ma, mb) and torch.equal(va, vb)


def test_kt_class_and_product_model(tmp_path):
    """scratch.coKriging's construction: kt = Kt(1, [0], var, lengthscale); kxy = nonDivK(2, [1,2], r);
    k = kt * kxy; GPRegression(X, obs, k) (scratch.py:506-511)."""
    X3, Xs3, y = _data(120, 40, 21)
    kt = myKernel.Kt(input_dim=1, active_dims=[0], var=5.0, lengthscale=1.0)
    # Kt alone: the tiled time RBF (myKernel.py:350-363)
    np.testing.assert_allclose(kt.K(X3[:, :1], Xs3[:, :1]), orc.kt_K(X3[:, 0], Xs3[:, 0], 5.0, 1.0), rtol=1e-13, atol=1e-15)
    assert kt.Kdiag(X3[:, :1]).shape == (240,) and np.all(kt.Kdiag(X3[:, :1]) == 5.0)
    k = kt * myKernel.nonDivK(2, [1, 2], 0.9)
    np.testing.assert_allclose(k.K(X3, Xs3), orc.st_K(X3, Xs3, 0.9, 1.0, 1.0, 5.0, 1.0), rtol=0, atol=1e-13)
    m = models.GPRegression(X3, y[:, None], k, noise_var=0.02)
    assert m.parameter_names() == ["mul.Kt.var", "mul.Kt.lengthscale", "mul.nonDivK.length", "Gaussian_noise.variance"]
    lo, go = orc.st_lml_and_grad(X3, y, 0.9, 1.0, 1.0, 5.0, 1.0, 0.02, jitter=1e-8)
    assert abs(m.log_likelihood() - lo) <= 1e-6 * abs(lo)
    np.testing.assert_allclose([p.gradient for p in m.parameters], [go[3], go[4], go[0], go[5]], rtol=1e-6, atol=1e-7)
    ll0 = m.log_likelihood()
    m.optimize(max_iters=30)
    assert m.log_likelihood() > ll0
    p = str(tmp_path / "st.pkl")
    m.pickle(p)
    m2 = models.load(p)
    np.testing.assert_array_equal(m2.param_array, m.param_array)
    a, b = m.predict(Xs3), m2.predict(Xs3)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    var, lt, ln, nz = m.param_array
    f = orc.st_fit(X3, y, ln, 1.0, 1.0, var, lt, nz, jitter=1e-8)
    mo, vo = orc.st_predict(X3, f, ln, 1.0, 1.0, var, lt, Xs3, var_add=nz)
    np.testing.assert_allclose(a[0][:, 0], mo, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(a[1][:, 0], vo, rtol=1e-7, atol=1e-10)
