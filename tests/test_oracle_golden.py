"""Pin the oracle (oracle/gp_oracle.py) against vectors produced by the reference's own
code (tests/golden/make_golden.py runs verbatim slices of the reference).  CPU only."""
import os

import numpy as np
import pytest

from oracle import gp_oracle as orc
from oracle import ref_slices as rs


@pytest.fixture(scope="module")
def ks(golden_dir):
    return np.load(os.path.join(golden_dir, "kernel_small.npz"))


def test_kernel_matches_three_reference_implementations(ks):
    X, X2 = ks["X"], ks["X2"]
    for t, (ldf, lcf, r) in enumerate(ks["thetas"]):
        K_sym = orc.helmholtz_K(X, None, ldf, lcf, r)
        K_x = orc.helmholtz_K(X, X2, ldf, lcf, r)
        # class form myKernel.py:27-53, function form GP_scripts.py:6-42, loops GP_scripts.py:57-123
        np.testing.assert_allclose(K_sym, ks["ref_K_class_sym_%d" % t], rtol=0, atol=1e-15)
        np.testing.assert_allclose(K_x, ks["ref_K_class_x_%d" % t], rtol=0, atol=1e-15)
        np.testing.assert_allclose(K_x, ks["ref_K_func_x_%d" % t], rtol=0, atol=1e-15)
        np.testing.assert_allclose(K_sym, ks["ref_K_loops_sym_%d" % t], rtol=0, atol=2e-15)
        # compute_Ks is K(X*, X): rows = grid
        np.testing.assert_allclose(orc.helmholtz_K(X2, X, ldf, lcf, r), ks["ref_Ks_loops_%d" % t],
                                   rtol=0, atol=2e-15)
        np.testing.assert_array_equal(orc.helmholtz_Kdiag(X2.shape[0], ldf, lcf, r),
                                      ks["ref_Kdiag_%d" % t])


def test_single_component_kernels(ks):
    X, X2 = ks["X"], ks["X2"]
    np.testing.assert_allclose(orc.helmholtz_K(X, X2, 1.7, 1.0, 1.0), ks["ref_nonDivK"], atol=1e-15)
    np.testing.assert_allclose(orc.helmholtz_K(X, X2, 1.0, 0.9, 0.0), ks["ref_nonRotK"], atol=1e-15)
    np.testing.assert_allclose(orc.helmholtz_Kdiag(X2.shape[0], 1.7, 1.0, 1.0), ks["ref_nonDivK_diag"])
    np.testing.assert_allclose(orc.helmholtz_Kdiag(X2.shape[0], 1.0, 0.9, 0.0), ks["ref_nonRotK_diag"])


def test_compat_gradient_reproduces_reference(ks):
    X, X2 = ks["X"], ks["X2"]
    for t, (ldf, lcf, r) in enumerate(ks["thetas"]):
        g = orc.kernel_grad_sums(ks["W_sym"], X, None, ldf, lcf, r, reference_compat=True)
        np.testing.assert_allclose(g, ks["ref_grad_compat_sym_%d" % t], rtol=1e-12, atol=1e-13)
        g = orc.kernel_grad_sums(ks["W_x"], X, X2, ldf, lcf, r, reference_compat=True)
        np.testing.assert_allclose(g, ks["ref_grad_compat_x_%d" % t], rtol=1e-12, atol=1e-13)
    g = orc.kernel_grad_sums(ks["W_x"], X, X2, 1.7, 1.0, 1.0, reference_compat=True)
    np.testing.assert_allclose(g[0], ks["ref_nonDivK_grad_compat"][0], rtol=1e-12)
    g = orc.kernel_grad_sums(ks["W_x"], X, X2, 1.0, 0.9, 0.0, reference_compat=True)
    np.testing.assert_allclose(g[1], ks["ref_nonRotK_grad_compat"][0], rtol=1e-12)


def test_correct_gradient_matches_central_differences(ks):
    X, X2, W = ks["X"], ks["X2"], ks["W_x"]
    h = 1e-6
    for (ldf, lcf, r) in [(2.0, 2.0, 0.5), (1.3, 3.1, 0.2)]:
        g = orc.kernel_grad_sums(W, X, X2, ldf, lcf, r)
        f = lambda a, b, c: np.sum(orc.helmholtz_K(X, X2, a, b, c) * W)
        fd = np.array([(f(ldf + h, lcf, r) - f(ldf - h, lcf, r)) / (2 * h),
                       (f(ldf, lcf + h, r) - f(ldf, lcf - h, r)) / (2 * h),
                       (f(ldf, lcf, r + h) - f(ldf, lcf, r - h)) / (2 * h)])
        np.testing.assert_allclose(g, fd, rtol=1e-7, atol=1e-8)
        # and the reference's formula is NOT the derivative
        gc = orc.kernel_grad_sums(W, X, X2, ldf, lcf, r, reference_compat=True)
        assert abs(gc[0] - fd[0]) > 1e-3 * abs(fd[0])
        np.testing.assert_allclose(gc[2], fd[2], rtol=1e-7)


def test_scalar_helpers_match_reference(ks):
    """GP_scripts.sqExp / rbf (GP_scripts.py:125-142) are instances of the scalar RBF family."""
    X, X2 = ks["X"], ks["X2"]
    np.testing.assert_allclose(orc.rbf_sum_K(X, X2, [1.0], [[1.7, 1.7]]), ks["ref_sqExp"], rtol=1e-14, atol=1e-16)
    np.testing.assert_allclose(orc.rbf_sum_K(X[:, :1], X2[:, :1], [0.64], [[1.3]]), ks["ref_rbf_x"], rtol=1e-14, atol=1e-16)
    np.testing.assert_allclose(orc.rbf_sum_K(X[:, :1], None, [0.64], [[1.3]]) + 0.05 * np.eye(X.shape[0]),
                               ks["ref_rbf_sym"], rtol=1e-14, atol=1e-16)


def test_known_answers():
    # K(0) = I / l^2; uu minimum -2 e^{-3/2} / l^2 at distance sqrt(3) l along y (div-free)
    l = 0.2
    K0 = orc.helmholtz_K(np.zeros((1, 2)), None, l, l, 1.0)
    np.testing.assert_allclose(K0, np.eye(2) / l ** 2)
    Xa = np.zeros((1, 2))
    Xb = np.array([[0.0, np.sqrt(3.0) * l]])
    Kdf = orc.helmholtz_K(Xa, Xb, l, l, 1.0)
    np.testing.assert_allclose(Kdf[0, 0], -2 * np.exp(-1.5) / l ** 2, rtol=1e-13)
    Xb = np.array([[np.sqrt(3.0) * l, 0.0]])
    Kcf = orc.helmholtz_K(Xa, Xb, l, l, 0.0)
    np.testing.assert_allclose(Kcf[0, 0], -2 * np.exp(-1.5) / l ** 2, rtol=1e-13)
    # uv sign: +d1 d2 (div-free), -d1 d2 (curl-free)
    Xb = np.array([[0.1, 0.1]])
    assert orc.helmholtz_K(Xa, Xb, l, l, 1.0)[0, 1] > 0
    assert orc.helmholtz_K(Xa, Xb, l, l, 0.0)[0, 1] < 0


@pytest.mark.parametrize("ts", [0, 100])
def test_simlaser_pipeline(golden_dir, ts):
    """Cholesky formulation (oracle.fit/predict) vs the reference's explicit-inverse numpy
    pipeline run through the reference's own functions (GP_laser.py:177-185,128-131)."""
    g = np.load(os.path.join(golden_dir, "simlaser_ts%d.npz" % ts))
    X, y, Xs = g["X"], g["y"], g["Xs"]
    ldf, lcf, r = g["theta"]
    noise = float(g["noise"])
    f = orc.fit(X, y, ldf, lcf, r, noise)
    mean, var = orc.predict(X, f, ldf, lcf, r, Xs)
    np.testing.assert_allclose(mean, g["ref_mean"], rtol=1e-8, atol=1e-10 * np.abs(g["ref_mean"]).max())
    np.testing.assert_allclose(var, g["ref_var"], rtol=1e-8)
    np.testing.assert_allclose(f["lml"], float(g["derived_lml"]), rtol=1e-12)
    # the simLaser K* weighting defect (GP_laser.py:181) reproduced on request
    m2, _ = orc.fit_predict_inverse_form(X, y, ldf, lcf, r, noise, Xs, ks_cf_weight=(1 - r) * r,
                                         want_var=False)
    np.testing.assert_allclose(m2, g["ref_mean_simlaser"], rtol=1e-9, atol=1e-12)
    if ts == 0:
        assert abs(f["lml"] - 317.0291777225) < 1e-9       # SURVEY.md §4 / BASELINE.md probe


def test_lml_grad_vs_finite_differences(golden_dir):
    g = np.load(os.path.join(golden_dir, "simlaser_ts0.npz"))
    X, y = g["X"][:120], np.concatenate([g["y"][:120], g["y"][400:520]])
    th = (1.3, 3.1, 0.2, 0.05)
    val, grad = orc.lml_and_grad(X, y, *th)
    h = 1e-6
    for p in range(4):
        a = list(th); b = list(th)
        a[p] += h; b[p] -= h
        fd = (orc.lml(X, y, *a) - orc.lml(X, y, *b)) / (2 * h)
        np.testing.assert_allclose(grad[p], fd, rtol=2e-6, atol=1e-6)


def test_generic_gp_algebra_vs_sklearn():
    """The fit/predict/LML algebra is sklearn's (krig.py:174-194 call site); check the same
    algebra on a scalar ARD-RBF + WhiteKernel against live scikit-learn."""
    from sklearn.gaussian_process import GaussianProcessRegressor, kernels
    import scipy.linalg as sla
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 10, size=(60, 3))
    yv = np.sin(X[:, 0]) + 0.1 * rng.normal(size=60)
    Xs = rng.uniform(0, 10, size=(25, 3))
    var0, ls, noise = 1.7, [1.1, 2.3, 0.9], 0.03
    k = var0 * kernels.RBF(length_scale=ls) + kernels.WhiteKernel(noise_level=noise)
    m = GaussianProcessRegressor(kernel=k, optimizer=None, alpha=0.0).fit(X, yv)
    mu_s, sd_s = m.predict(Xs, return_std=True)
    K = orc.rbf_ard_K(X, None, var0, ls) + noise * np.eye(60)
    L = sla.cholesky(K, lower=True)
    alpha = sla.cho_solve((L, True), yv)
    Ks = orc.rbf_ard_K(Xs, X, var0, ls)
    V = sla.solve_triangular(L, Ks.T, lower=True)
    var = var0 + noise - np.einsum("ij,ij->j", V, V)          # WhiteKernel is in sklearn's diag
    np.testing.assert_allclose(Ks @ alpha, mu_s, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(np.sqrt(var), sd_s, rtol=1e-8)
    lml = -0.5 * yv @ alpha - np.log(np.diag(L)).sum() - 30 * np.log(2 * np.pi)
    np.testing.assert_allclose(lml, m.log_marginal_likelihood_value_, rtol=1e-10)


@pytest.mark.skipif(not rs.available(), reason="reference tree not present (GPU box)")
def test_live_reference_slices_agree_with_oracle():
    rng = np.random.default_rng(11)
    X = rng.uniform(0, 5, size=(13, 2))
    X2 = rng.uniform(0, 5, size=(7, 2))
    k = rs.mykernel_class(1.1, 2.7, 0.35)
    np.testing.assert_allclose(orc.helmholtz_K(X, X2, 1.1, 2.7, 0.35), k.K(X, X2), atol=1e-15)
    gs = rs.gp_scripts()
    np.testing.assert_allclose(orc.helmholtz_K(X, X2, 1.1, 2.7, 0.35),
                               gs["myKernel"](X, X2, 1.1, 2.7, 0.35), atol=1e-15)


# ---- scalar ARD-RBF sum family: oracle pinned against live scikit-learn (krig.py:174-194) --------
def _hp_split(HP):
    """HP as krig.scikit_prior reads a GPy param_array: [var, l_t, l_y, l_x](, [var, l_t, l_y, l_x]), noise."""
    Q = (HP.size - 1) // 4
    var = [HP[4 * q] for q in range(Q)]
    ls = [HP[4 * q + 1:4 * q + 4] for q in range(Q)]
    return var, ls, float(HP[-1])


@pytest.mark.parametrize("name", ["sklearn_rbf1", "sklearn_rbf2"])
def test_rbf_oracle_matches_scikit_learn(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    var, ls, noise = _hp_split(g["HP"])
    XT, u, Xg = g["XT"], g["u"], g["Xg"]
    np.testing.assert_allclose(orc.rbf_sum_K(XT[:40], Xg[:30], var, ls), g["ref_K_signal"], rtol=1e-13, atol=1e-16)
    Ktr = orc.rbf_sum_K(XT[:50], None, var, ls) + noise * np.eye(50)
    np.testing.assert_allclose(Ktr, g["ref_K_train"], rtol=1e-13, atol=1e-16)
    f = orc.rbf_fit(XT, u, var, ls, noise, jitter=float(g["sklearn_alpha"]))
    assert abs(f["lml"] - float(g["ref_lml"])) <= 1e-9 * abs(float(g["ref_lml"]))
    mean, v = orc.rbf_predict(XT, f, var, ls, Xg, var_add=noise)       # WhiteKernel is in sklearn's variance
    np.testing.assert_allclose(mean, g["ref_mean"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(v, g["ref_var"], rtol=1e-8)
    # sklearn differentiates w.r.t. log(theta): d/dlog(theta) = theta * d/dtheta
    lml, grad = orc.rbf_lml_and_grad(XT, u, var, ls, noise, jitter=float(g["sklearn_alpha"]))
    theta = np.exp(g["ref_theta"])
    np.testing.assert_allclose(grad * theta, g["ref_grad_logtheta"], rtol=1e-7, atol=1e-8)


def test_rbf_kernel_gradient_matches_central_differences(golden_dir):
    g = np.load(os.path.join(golden_dir, "sklearn_rbf2.npz"))
    var, ls, _ = _hp_split(g["HP"])
    X, X2 = g["XT"][:30], g["Xg"][:20]
    W = np.random.default_rng(1).normal(size=(30, 20))
    got = orc.rbf_kernel_grad_sums(W, X, X2, var, ls)
    theta = np.concatenate([[var[q]] + list(ls[q]) for q in range(2)])

    def f(th):
        return np.sum(orc.rbf_sum_K(X, X2, [th[0], th[4]], [th[1:4], th[5:8]]) * W)
    fd = np.array([(f(theta + 1e-6 * e) - f(theta - 1e-6 * e)) / 2e-6 for e in np.eye(8)])
    np.testing.assert_allclose(got, fd, rtol=1e-6, atol=1e-8)


# ---- space-time product kernel (formula restatement; the upstream path is dead code) --------------
def test_spacetime_oracle_consistency():
    rng = np.random.default_rng(5)
    X3 = np.stack([rng.uniform(0, 6, 12), rng.uniform(0, 10, 12), rng.uniform(0, 10, 12)], axis=1)
    Xs3 = np.stack([rng.uniform(0, 6, 7), rng.uniform(0, 10, 7), rng.uniform(0, 10, 7)], axis=1)
    W = rng.normal(size=(24, 14))
    theta = np.array([1.3, 3.1, 0.2, 0.7, 2.5])
    go = np.array([np.sum(d * W) for d in orc.st_dK(X3, Xs3, *theta)])
    f = lambda th: np.sum(orc.st_K(X3, Xs3, *th) * W)
    fd = np.array([(f(theta + 1e-6 * e) - f(theta - 1e-6 * e)) / 2e-6 for e in np.eye(5)])
    np.testing.assert_allclose(go, fd, rtol=1e-7, atol=1e-8)
    # Kt.K is the time RBF tiled 2x2 (myKernel.py:357-358); the product is elementwise (GPy Prod)
    C = orc.rbf_ard_K(X3[:, :1], Xs3[:, :1], 0.7, [2.5])
    np.testing.assert_array_equal(orc.kt_K(X3[:, 0], Xs3[:, 0], 0.7, 2.5), np.tile(C, (2, 2)))
    np.testing.assert_array_equal(orc.st_K(X3, Xs3, *theta), np.tile(C, (2, 2)) * orc.helmholtz_K(X3[:, 1:], Xs3[:, 1:], 1.3, 3.1, 0.2))
    # equal times, unit variance: exactly the pinned Helmholtz kernel
    X0 = X3.copy()
    X0[:, 0] = 1.0
    np.testing.assert_array_equal(orc.st_K(X0, None, 1.3, 3.1, 0.2, 1.0, 9.0), orc.helmholtz_K(X0[:, 1:], None, 1.3, 3.1, 0.2))
    y = rng.normal(size=24)
    lo, g = orc.st_lml_and_grad(X3, y, *theta, 0.1)
    h = 1e-6
    for i in range(5):
        e = np.zeros(5); e[i] = h
        fdl = (orc.st_fit(X3, y, *(theta + e), 0.1)["lml"] - orc.st_fit(X3, y, *(theta - e), 0.1)["lml"]) / (2 * h)
        assert abs(g[i] - fdl) <= 1e-5 * max(1.0, abs(fdl))


# ---- sum of space-time Helmholtz terms (krig.py:396-407; module myKernel2 missing upstream) ----------
def test_hsum_oracle_reduces_to_the_pinned_kernels():
    """The term-sum restatement has no reference code to be pinned against, but its isotropic
    subspace is the golden-checked Helmholtz kernel (myKernel.py:27-53) and, with a shared time
    scale, the space-time product (scratch.py:506-508)."""
    rng = np.random.default_rng(11)
    X, X2 = rng.uniform(0, 8, (23, 2)), rng.uniform(0, 8, (17, 2))
    for (l_df, l_cf, ratio) in [(2.0, 2.0, 0.5), (1.3, 3.1, 0.2), (0.8, 1.9, 1.0), (0.8, 1.9, 0.0)]:
        types, params = [0, 1], [[ratio, 1.0, l_df, l_df], [1 - ratio, 1.0, l_cf, l_cf]]
        np.testing.assert_allclose(orc.hsum_K(X, X2, types, params), orc.helmholtz_K(X, X2, l_df, l_cf, ratio), rtol=0, atol=2e-15)
        np.testing.assert_allclose(orc.hsum_Kdiag(5, types, params), orc.helmholtz_Kdiag(5, l_df, l_cf, ratio), rtol=1e-15)
    X3, X3b = np.c_[rng.uniform(0, 3, 23), X], np.c_[rng.uniform(0, 3, 17), X2]
    params = [[0.2 * 1.7, 0.9, 1.3, 1.3], [0.8 * 1.7, 0.9, 3.1, 3.1]]
    np.testing.assert_allclose(orc.hsum_K(X3, X3b, [0, 1], params), orc.st_K(X3, X3b, 1.3, 3.1, 0.2, 1.7, 0.9), rtol=0, atol=2e-15)


def test_hsum_oracle_gradient_and_validity():
    rng = np.random.default_rng(12)
    X3, X3b = rng.uniform(0, 6, (14, 3)), rng.uniform(0, 6, (9, 3))
    types = [0, 1, 0]
    P = np.array([[0.8, 1.1, 1.3, 0.7], [1.2, 0.6, 2.1, 1.5], [0.5, 2.0, 0.9, 1.9]])
    d = orc.hsum_dK(X3, X3b, types, P)
    for q in range(3):
        for j in range(4):
            Pp, Pm = P.copy(), P.copy()
            Pp[q, j] += 1e-6
            Pm[q, j] -= 1e-6
            fd = (orc.hsum_K(X3, X3b, types, Pp) - orc.hsum_K(X3, X3b, types, Pm)) / 2e-6
            np.testing.assert_allclose(d[q][j], fd, rtol=0, atol=2e-9 * max(1.0, np.abs(fd).max()))
    # a valid covariance: symmetric, positive semi-definite, diagonal = Kdiag
    K = orc.hsum_K(X3, None, types, P)
    np.testing.assert_array_equal(K, K.T)
    assert np.linalg.eigvalsh(K).min() > -1e-12
    np.testing.assert_allclose(np.diag(K), orc.hsum_Kdiag(14, types, P), rtol=1e-14)
    # the LML gradient is the derivative of the LML
    y = rng.normal(size=28)
    l0, g0 = orc.hsum_lml_and_grad(X3, y, types, P, 0.1)
    for q, j in [(0, 0), (1, 1), (2, 2), (1, 3)]:
        Pp, Pm = P.copy(), P.copy()
        Pp[q, j] += 1e-6
        Pm[q, j] -= 1e-6
        fd = (orc.hsum_fit(X3, y, types, Pp, 0.1)["lml"] - orc.hsum_fit(X3, y, types, Pm, 0.1)["lml"]) / 2e-6
        assert abs(fd - g0[4 * q + j]) <= 1e-6 * max(1.0, abs(fd))
    fd = (orc.hsum_fit(X3, y, types, P, 0.1 + 1e-6)["lml"] - orc.hsum_fit(X3, y, types, P, 0.1 - 1e-6)["lml"]) / 2e-6
    assert abs(fd - g0[-1]) <= 1e-6 * max(1.0, abs(fd))


def test_morton_order_is_a_locality_preserving_permutation():
    """oracle.morton_order (the restatement of csrc/order.cu): a permutation, stable on ties, invariant under
    translation and uniform scaling, and consecutive points are close: groups of 16 consecutive points of a jittered
    lattice span a few lattice cells, not the domain."""
    rng = np.random.default_rng(2)
    gx, gy = np.meshgrid(np.arange(40) * 0.5, np.arange(40) * 0.5)
    P = np.stack([gx.ravel(), gy.ravel()], 1) + rng.uniform(-0.2, 0.2, (1600, 2))
    P = P[rng.permutation(1600)]
    perm = orc.morton_order(P)
    assert sorted(perm.tolist()) == list(range(1600))
    np.testing.assert_array_equal(orc.morton_order(3.0 * P + 11.0), perm)
    Q = P[perm]
    spans = [np.ptp(Q[i:i + 16], axis=0).max() for i in range(0, 1600, 16)]
    assert np.median(spans) < 3.0 and np.ptp(P, axis=0).max() > 19.0
    D = np.array([[0.0, 0.0], [1.0, 1.0], [0.0, 0.0], [1.0, 1.0]])
    np.testing.assert_array_equal(orc.morton_order(D), [0, 2, 1, 3])

