#!/usr/bin/env python
"""Golden vectors for the scalar ARD-RBF path from LIVE scikit-learn, built exactly as the
reference's krig.scikit_prior does (krig.py:174-194):

    k = HP[0]*RBF(length_scale=[HP1,HP2,HP3]) (+ HP[4]*RBF([HP5,HP6,HP7])) + WhiteKernel(noise)
    GaussianProcessRegressor(kernel=k, optimizer=None).fit(XT, u).predict(X, return_std=True)

Run in the build container:  python tests/golden/make_golden_sklearn.py
Every ``ref_*`` array comes from scikit-learn (version recorded), not from the oracle."""
import os

import numpy as np
import sklearn
from sklearn.gaussian_process import GaussianProcessRegressor, kernels

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, HP, N=180, M=90, seed=11):
    rng = np.random.default_rng(seed)
    # (t, y, x) rows like krig: hours, km, km
    XT = np.stack([rng.uniform(0, 12, N), rng.uniform(0, 15, N), rng.uniform(-5, 15, N)], axis=1)
    u = (0.3 * np.sin(XT[:, 1] / 3.0) * np.cos(XT[:, 2] / 4.0) + 0.05 * XT[:, 0] / 12.0
         + rng.normal(0, 0.03, N))[:, None]
    Xg = np.stack([np.full(M, 6.0), rng.uniform(1, 15, M), rng.uniform(-5, 15, M)], axis=1)
    noise = HP[-1]
    k = HP[0] * kernels.RBF(length_scale=[HP[1], HP[2], HP[3]])
    if HP.size - 1 > 5:
        k = k + HP[4] * kernels.RBF(length_scale=[HP[5], HP[6], HP[7]])
    k = k + kernels.WhiteKernel(noise_level=noise)
    model = GaussianProcessRegressor(kernel=k, optimizer=None)
    model.fit(XT, u)
    U, Ustd = model.predict(Xg, return_std=True)
    lml, grad = model.log_marginal_likelihood(model.kernel_.theta, eval_gradient=True)
    # K of the signal part only (what gp2d_rbf_kernel_build computes)
    ksig = model.kernel_.k1
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), HP=HP, XT=XT, u=u[:, 0], Xg=Xg,
        ref_mean=np.reshape(U, [-1]), ref_var=np.reshape(Ustd, [-1]) ** 2, ref_lml=lml,
        ref_grad_logtheta=grad, ref_theta=model.kernel_.theta,
        ref_K_signal=ksig(XT[:40], Xg[:30]), ref_K_train=model.kernel_(XT[:50]),
        sklearn_version=sklearn.__version__, sklearn_alpha=model.alpha)
    print(name, "lml", lml, "var range", (Ustd ** 2).min(), (Ustd ** 2).max())


if __name__ == "__main__":
    case("sklearn_rbf1", np.array([0.09, 8.0, 4.0, 5.0, 0.0009]))
    case("sklearn_rbf2", np.array([0.05, 20.0, 6.0, 7.0, 0.02, 3.0, 1.5, 2.0, 0.0009]), seed=12)


def optimised_case():
    """scikit-learn's default behaviour (optimizer='fmin_l_bfgs_b'; testKrig.py:139-140): the
    hyper-parameters are fitted by maximising the log-marginal likelihood over log(theta)."""
    rng = np.random.default_rng(21)
    N, M = 150, 60
    XT = np.stack([rng.uniform(0, 12, N), rng.uniform(0, 15, N), rng.uniform(-5, 15, N)], axis=1)
    u = (0.3 * np.sin(XT[:, 1] / 3.0) * np.cos(XT[:, 2] / 4.0) + 0.05 * XT[:, 0] / 12.0 + rng.normal(0, 0.03, N))
    Xg = np.stack([np.full(M, 6.0), rng.uniform(1, 15, M), rng.uniform(-5, 15, M)], axis=1)
    k = 0.1 * kernels.RBF(length_scale=[5.0, 3.0, 3.0]) + kernels.WhiteKernel(noise_level=0.01)
    m = GaussianProcessRegressor(kernel=k, n_restarts_optimizer=0).fit(XT, u)
    U, Ustd = m.predict(Xg, return_std=True)
    np.savez_compressed(os.path.join(HERE, "sklearn_rbf_opt.npz"), XT=XT, u=u, Xg=Xg,
                        start=np.array([0.1, 5.0, 3.0, 3.0, 0.01]), ref_theta=m.kernel_.theta,
                        ref_lml=m.log_marginal_likelihood_value_, ref_mean=U, ref_var=Ustd ** 2,
                        sklearn_version=sklearn.__version__)
    print("sklearn_rbf_opt", m.kernel_, m.log_marginal_likelihood_value_)


if __name__ == "__main__":
    optimised_case()
