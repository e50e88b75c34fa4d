"""GPU tests of the sum-of-space-time-Helmholtz-terms family (what
krig.kriging(kernelType=2,3,4,nKernels) asks the missing module myKernel2 for, krig.py:396-407)
against the CPU oracle.  On its isotropic subspace the family must reproduce the pinned Helmholtz
and space-time entry points; the anisotropic part is specified by the oracle ("parity unpinned")."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import gp2d_b200 as gp                                        # noqa: E402
from oracle import gp_oracle as orc                          # noqa: E402

# (types, params[Q,4] = var, lt, la, lb)
CASES = [
    ([0], [[1.0, 1.0, 1.0, 1.0]]),                                            # krig.py:397 defaults
    ([1], [[0.7, 2.5, 1.3, 0.8]]),
    ([0, 1], [[0.6, 1.5, 1.3, 2.0], [1.4, 0.7, 3.1, 2.2]]),                    # kernelType 4
    ([0, 1, 0, 1], [[0.6, 1.5, 1.3, 2.0], [1.4, 0.7, 3.1, 2.2], [0.2, 4.0, 0.6, 0.5], [0.9, 2.2, 5.0, 4.0]]),
    ([0, 1, 1, 0, 0, 1, 0, 1], [[0.3 + 0.1 * q, 1.0 + 0.3 * q, 0.8 + 0.4 * q, 2.9 - 0.3 * q] for q in range(8)]),
]


def _data(N, M, seed, ldx=3):
    rng = np.random.default_rng(seed)
    X = np.stack([rng.uniform(0, 6, N), rng.uniform(0, 10, N), rng.uniform(0, 10, N)], axis=1)
    Xs = np.stack([rng.uniform(0, 6, M), rng.uniform(0, 10, M), rng.uniform(0, 10, M)], axis=1)
    y = np.concatenate([np.sin(X[:, 1] / 2) * np.cos(X[:, 0] / 3), np.cos(X[:, 2] / 2)]) + 0.05 * rng.normal(size=2 * N)
    if ldx == 2:
        X, Xs = X[:, 1:], Xs[:, 1:]
    return np.ascontiguousarray(X), np.ascontiguousarray(Xs), y


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("N,M,ldx", [(1, 1, 3), (9, 6, 2), (130, 257, 3), (65, 128, 2)])
def test_kernel_kdiag_and_gradient_sums(case, N, M, ldx):
    types, params = case
    X, Xs, _ = _data(N, M, N + M, ldx)
    scale = 4 * sum(p[0] / min(p[2], p[3]) ** 2 for p in params)
    np.testing.assert_allclose(gp.hsum_K(X, Xs, types, params).cpu().numpy(), orc.hsum_K(X, Xs, types, params), rtol=0, atol=1e-14 * scale)
    K = gp.hsum_K(X, None, types, params, diag_add=0.3).cpu().numpy()
    np.testing.assert_allclose(K, orc.hsum_K(X, None, types, params) + 0.3 * np.eye(2 * N), rtol=0, atol=1e-14 * scale)
    np.testing.assert_allclose(gp.hsum_Kdiag(M, ldx, types, params).cpu().numpy(), orc.hsum_Kdiag(M, types, params), rtol=1e-15)
    np.testing.assert_allclose(np.diag(K) - 0.3, orc.hsum_Kdiag(N, types, params), rtol=1e-13)
    W = np.random.default_rng(3).normal(size=(2 * N, 2 * M))
    g = gp.hsum_grad_sums(W, X, Xs, types, params).cpu().numpy()
    go = orc.hsum_kernel_grad_sums(W, X, Xs, types, params)
    np.testing.assert_allclose(g, go, rtol=1e-10, atol=1e-11 * scale * max(1, np.sqrt(N * M)))


def test_isotropic_subspace_is_the_reference_kernel():
    """la == lb, no time, {div-free var = ratio} + {curl-free var = 1 - ratio} = myKernel.K
    (myKernel.py:27-53); with a shared time scale = the space-time product kernel."""
    X, Xs, y = _data(150, 211, 4, ldx=2)
    l_df, l_cf, ratio = 1.3, 3.1, 0.2
    types, params = [0, 1], [[ratio, 1.0, l_df, l_df], [1 - ratio, 1.0, l_cf, l_cf]]
    np.testing.assert_allclose(gp.hsum_K(X, Xs, types, params).cpu().numpy(), gp.kernel_K(X, Xs, l_df, l_cf, ratio).cpu().numpy(),
                               rtol=0, atol=1e-15)
    a = gp.HelmholtzSumGP(X, y, types, params, 0.05)
    b = gp.HelmholtzGP(X, y, l_df, l_cf, ratio, 0.05)
    la, ga = a.lml_and_grad()
    lb, gb = b.lml_and_grad()
    assert abs(la - lb) <= 1e-10 * abs(lb)
    # chain rule onto (l_df, l_cf, ratio, noise): dl = d/dla + d/dlb, dratio = d/dvar_0 - d/dvar_1
    ga = ga.reshape(-1)
    np.testing.assert_allclose([ga[2] + ga[3], ga[6] + ga[7], ga[0] - ga[4], ga[8]], gb, rtol=1e-8, atol=1e-9)
    ma, va = a.predict(Xs)
    mb, vb = b.predict(Xs)
    np.testing.assert_allclose(ma.cpu().numpy(), mb.cpu().numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(va.cpu().numpy(), vb.cpu().numpy(), rtol=1e-10, atol=1e-13)
    X3, Xs3, y = _data(150, 211, 4, ldx=3)
    tvar, lt = 1.7, 0.9
    params3 = [[ratio * tvar, lt, l_df, l_df], [(1 - ratio) * tvar, lt, l_cf, l_cf]]
    np.testing.assert_allclose(gp.hsum_K(X3, Xs3, types, params3).cpu().numpy(),
                               gp.st_K(X3, Xs3, l_df, l_cf, ratio, tvar, lt).cpu().numpy(), rtol=0, atol=1e-15)


@pytest.mark.parametrize("N,M,ldx,case", [(1, 3, 3, CASES[0]), (65, 130, 2, CASES[1]), (300, 517, 3, CASES[2]),
                                          (700, 2601, 3, CASES[3]), (257, 64, 2, CASES[4]), (200, 100, 3, CASES[4])])
def test_fit_predict_lml_grad_vs_oracle(N, M, ldx, case):
    types, params = case
    X, Xs, y = _data(N, M, 100 + N, ldx)
    noise = 0.02
    g = gp.HelmholtzSumGP(X, y, types, params, noise, jitter=1e-8)
    lml, grad = g.lml_and_grad()
    mean, var = g.predict(Xs, include_noise=True)
    f = orc.hsum_fit(X, y, types, params, noise, jitter=1e-8)
    mo, vo = orc.hsum_predict(X, f, types, params, Xs, var_add=noise)
    lo, go = orc.hsum_lml_and_grad(X, y, types, params, noise, jitter=1e-8)
    assert abs(lml - lo) <= 1e-6 * max(abs(lo), 1.0)
    np.testing.assert_allclose(grad, go, rtol=1e-6, atol=1e-7 * max(1.0, np.abs(go).max()))
    np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-8, atol=1e-9 * max(np.abs(mo).max(), 1e-3))
    np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(g.alpha().cpu().numpy(), f["alpha"], rtol=1e-7, atol=1e-8 * np.abs(f["alpha"]).max())
    if M > 300:
        # any sub-grid gives the same bits (fixed row-block grouping)
        m1, v1 = g.predict(Xs[:300], include_noise=True)
        assert torch.equal(m1, torch.cat([mean[:300], mean[M:M + 300]])) and torch.equal(v1, torch.cat([var[:300], var[M:M + 300]]))


def test_anisotropic_prior_variances_differ_per_component():
    """Kdiag is (var/lb^2, var/la^2) for a divergence-free term: far from the data the predictive
    variance of the two components must tend to those two different values."""
    X, _, y = _data(50, 1, 8, ldx=2)
    types, params = [0], [[2.0, 1.0, 0.5, 2.0]]
    g = gp.HelmholtzSumGP(X, y, types, params, 0.05)
    far = np.array([[1e3, 1e3], [-1e3, 2e3]])
    mean, var = g.predict(far)
    np.testing.assert_allclose(var.cpu().numpy(), [2.0 / 4.0, 2.0 / 4.0, 2.0 / 0.25, 2.0 / 0.25], rtol=1e-12)
    assert float(mean.abs().max()) < 1e-12


def test_argument_checks():
    X, Xs, y = _data(10, 4, 1, ldx=3)
    with pytest.raises(ValueError):
        gp.hsum_K(X, Xs, [2], [[1, 1, 1, 1]])
    with pytest.raises(ValueError):
        gp.hsum_K(X, Xs, [0] * 9, [[1, 1, 1, 1]] * 9)
    with pytest.raises(ValueError):
        gp.hsum_K(X, Xs[:, 1:], [0], [[1, 1, 1, 1]])
    from gp2d_b200._lib import Gp2dError
    with pytest.raises(Gp2dError):
        gp.hsum_K(X, Xs, [0], [[1, 1, -1, 1]])
    with pytest.raises(Gp2dError):
        gp.hsum_K(X, Xs, [0], [[1, 0.0, 1, 1]])          # lt must be positive with time inputs
    gp.hsum_K(X[:, 1:], Xs[:, 1:], [0], [[1, 0.0, 1, 1]])  # ... and is ignored without
    g = gp.HelmholtzSumGP(X, y, [0, 1], [[1, 1, 1, 1], [1, 1, 2, 2]], 0.1)
    with pytest.raises(ValueError):
        g.set_params([0], [[1, 1, 1, 1]], 0.1)
    # not positive definite: zero variance, zero noise
    g = gp.HelmholtzSumGP(X, y, [0], [[0.0, 1, 1, 1]], 0.0)
    with pytest.raises(gp.LinAlgError):
        g.fit()


# ---- the reference-facing surface: myKernel2.divFreeK / curlFreeK and krig.kriging(kernelType=2..4) -------
def test_mykernel2_classes_and_model(tmp_path):
    """krig.py:396-411: k2 = divFreeK(input_dim=3) + curlFreeK(input_dim=3); k = k2 + k2; GPRegression(X, obs, k)."""
    from gp2d_b200 import models, myKernel2
    X, Xs, y = _data(150, 60, 31, ldx=3)
    kd = myKernel2.divFreeK(input_dim=3, active_dims=[0, 1, 2], var=1., lt=1., ly=1., lx=1.)      # krig.py:397 verbatim
    assert [p.name for p in kd.parameters] == ["var", "lt", "ly", "lx"] and kd.parameter_names()[0] == "divFreeK.var"
    np.testing.assert_allclose(kd.K(X, Xs), orc.hsum_K(X, Xs, [0], [[1, 1, 1, 1]]), rtol=0, atol=1e-13)
    np.testing.assert_allclose(kd.Kdiag(Xs), orc.hsum_Kdiag(60, [0], [[1, 1, 1, 1]]))
    k2 = myKernel2.divFreeK(input_dim=3, var=0.6, lt=1.5, ly=1.3, lx=2.0) + myKernel2.curlFreeK(input_dim=3, var=1.4, lt=0.7, ly=3.1, lx=2.2)
    k = k2.copy()
    k = k + k2                                                                                    # nKernels = 2 (krig.py:405-407)
    assert len(k.terms_list()) == 4 and len(k.parameters) == 16
    assert k.parameter_names()[:5] == ["sum.divFreeK.var", "sum.divFreeK.lt", "sum.divFreeK.ly", "sum.divFreeK.lx", "sum.curlFreeK.var"]
    assert k.parameter_names()[8] == "sum.divFreeK_1.var"
    # the copies are independent parameters
    k.terms_list()[2].ly.value = 0.6
    assert float(k.terms_list()[0].ly) == 1.3
    types, params = k.hsum_params()
    m = models.GPRegression(X, y[:, None], k, noise_var=0.02)
    lo, go = orc.hsum_lml_and_grad(X, y, types, params, 0.02, jitter=1e-8)
    assert abs(m.log_likelihood() - lo) <= 1e-6 * abs(lo)
    np.testing.assert_allclose([p.gradient for p in m.parameters], go, rtol=1e-6, atol=1e-7)
    # plug-in gradient protocol (what GPy would call)
    W = np.random.default_rng(0).normal(size=(300, 120))
    k.update_gradients_full(W, X, Xs)
    np.testing.assert_allclose([p.gradient for p in k.parameters], orc.hsum_kernel_grad_sums(W, X, Xs, types, params).reshape(-1),
                               rtol=1e-10, atol=1e-10)
    ll0 = m.log_likelihood()
    m.optimize(max_iters=25)
    assert m.log_likelihood() > ll0
    path = str(tmp_path / "k2.pkl")
    m.pickle(path)
    m2 = models.load(path)
    np.testing.assert_array_equal(m2.param_array, m.param_array)
    assert m2.parameter_names() == m.parameter_names()
    a, b = m.predict(Xs), m2.predict(Xs)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    types, params = m.kern.hsum_params()
    nz = float(m.Gaussian_noise)
    f = orc.hsum_fit(X, y, types, params, nz, jitter=1e-8)
    mo, vo = orc.hsum_predict(X, f, types, params, Xs, var_add=nz)
    np.testing.assert_allclose(a[0][:, 0], mo, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(a[1][:, 0], vo, rtol=1e-7, atol=1e-10)
    # restarts, sequential and concurrent, agree
    m3 = models.GPRegression(X, y[:, None], k2.copy(), noise_var=0.02)      # the model optimises its kernel in place
    m3.optimize_restarts(num_restarts=3, verbose=False, max_iters=15, seed=5)
    m4 = models.GPRegression(X, y[:, None], k2.copy(), noise_var=0.02)
    m4.optimize_restarts(num_restarts=3, verbose=False, max_iters=15, seed=5, parallel=3)
    np.testing.assert_allclose(m3.param_array, m4.param_array, rtol=1e-9)


def test_mykernel2_two_dimensional_and_column_selection():
    from gp2d_b200 import models, myKernel2
    X, Xs, y = _data(80, 30, 33, ldx=3)
    # input_dim=2 on columns (1, 2) of the (t, y, x) rows: no time factor, 3 parameters per term
    k = myKernel2.curlFreeK(input_dim=2, active_dims=[1, 2], var=0.8, ly=1.2, lx=2.5)
    assert [p.name for p in k.parameters] == ["var", "ly", "lx"]
    np.testing.assert_allclose(k.K(X, Xs), orc.hsum_K(X[:, 1:], Xs[:, 1:], [1], [[0.8, 1.0, 1.2, 2.5]]), rtol=0, atol=1e-13)
    m = models.GPRegression(X, y[:, None], k, noise_var=0.05)
    lo, go = orc.hsum_lml_and_grad(X[:, 1:], y, [1], [[0.8, 1.0, 1.2, 2.5]], 0.05, jitter=1e-8)
    assert abs(m.log_likelihood() - lo) <= 1e-6 * abs(lo)
    np.testing.assert_allclose([p.gradient for p in m.parameters], [go[0], go[2], go[3], go[4]], rtol=1e-6, atol=1e-7)
    m.optimize_restarts(num_restarts=2, verbose=False, max_iters=10, seed=1, parallel=2)
    mean, var = m.predict(Xs)
    assert mean.shape == (60, 1) and np.all(var > 0)
    with pytest.raises(ValueError):
        myKernel2.divFreeK(input_dim=3) + myKernel2.divFreeK(input_dim=2)
    with pytest.raises(ValueError):
        myKernel2.HelmholtzSum([myKernel2.divFreeK()] * 9)


def test_hsum_large_size_properties():
    """N = 8192 observations (16384 x 16384 covariance), four space-time terms: the oracle cannot
    factorise this in test time, so parity rests on size-independent properties -- the GP identity
    K alpha = y - noise alpha at observation sites (ties build, Cholesky, inverse, alpha and the
    fused predictive mean together), variance bounds, the far-field prior variances per component,
    grid-partition invariance and agreement of the two likelihood entry points."""
    from gp2d_b200 import synthetic
    N = 8192
    X2, y = synthetic.drifter_snapshot(N, config_id=5)
    rng = np.random.default_rng(1)
    X = np.ascontiguousarray(np.c_[rng.uniform(0, 2, N), X2])
    types, params = CASES[3]
    noise = 0.05
    g = gp.HelmholtzSumGP(X, y, types, params, noise)
    lml = g.fit()
    assert np.isfinite(lml)
    sel = rng.choice(N, 400, replace=False)
    pm, pv = g.predict(X[sel])
    al = g.alpha()
    idx = torch.as_tensor(sel, device=al.device)
    rhs = torch.cat([g.y[idx], g.y[N + idx]]) - noise * torch.cat([al[idx], al[N + idx]])
    torch.testing.assert_close(pm, rhs, rtol=1e-9, atol=1e-11)
    assert float(pv.min()) > 0.0 and float(pv.max()) < noise
    far = np.array([[1.0, 1e4, 1e4], [0.5, -1e4, 3e4]])
    fm, fv = g.predict(far)
    kd = orc.hsum_Kdiag(2, types, params)
    np.testing.assert_allclose(fv.cpu().numpy(), kd, rtol=1e-12)
    assert float(fm.abs().max()) < 1e-12
    G = np.ascontiguousarray(np.c_[np.full(600, 1.0), synthetic.prediction_grid(X2, 30, 20)])
    ma, va = g.predict(G)
    m1, v1 = g.predict(G[:217])
    m2, v2 = g.predict(G[217:])
    assert torch.equal(torch.cat([m1[:217], m2[:383], m1[217:], m2[383:]]), ma)
    assert torch.equal(torch.cat([v1[:217], v2[:383], v1[217:], v2[383:]]), va)
    lml2, grad = g.lml_and_grad()
    assert lml2 == lml and np.all(np.isfinite(grad)) and grad.shape == (17,)
