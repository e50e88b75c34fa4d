#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference itself.

Run in the build container (needs /root/reference; nothing here runs on the GPU box):

    python tests/golden/make_golden.py

Every array named ``ref_*`` is produced by *verbatim line slices of the reference*
(oracle/ref_slices.py) -- not by the oracle restatement -- so the fixtures pin the
oracle rather than echo it.  Arrays named ``derived_*`` combine reference-built
matrices with scipy LAPACK (the reference has no code for that quantity).
"""
import os
import sys
import time

import numpy as np
import scipy.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

os.environ["GP2D_EXEC_REFERENCE"] = "1"      # this tool exists to run the reference's own lines: explicit opt-in
from oracle import ref_slices as rs          # noqa: E402
from oracle import gp_oracle as orc          # noqa: E402  (only for haversine/host prep)

THETAS = [(2.0, 2.0, 0.5), (1.3, 3.1, 0.2), (0.6, 0.6, 1.0), (0.7, 1.9, 0.0)]


def kernel_small():
    rng = np.random.default_rng(20160207)
    X = rng.uniform(0, 6, size=(9, 2))
    X2 = rng.uniform(0, 6, size=(6, 2))
    W_sym = rng.normal(size=(18, 18))
    W_sym = 0.5 * (W_sym + W_sym.T)
    W_x = rng.normal(size=(18, 12))
    gs = rs.gp_scripts()
    out = {"X": X, "X2": X2, "thetas": np.array(THETAS), "W_sym": W_sym, "W_x": W_x}
    for t, (ldf, lcf, r) in enumerate(THETAS):
        k = rs.mykernel_class(ldf, lcf, r)
        out["ref_K_class_sym_%d" % t] = k.K(X, None)
        out["ref_K_class_x_%d" % t] = k.K(X, X2)
        out["ref_Kdiag_%d" % t] = k.Kdiag(X2)
        out["ref_K_func_x_%d" % t] = gs["myKernel"](X, X2, ldf, lcf, r)
        out["ref_K_loops_sym_%d" % t] = (r * gs["compute_K"](X[:, 0], X[:, 1], ldf, 1)
                                         + (1 - r) * gs["compute_K"](X[:, 0], X[:, 1], lcf, 2))
        # compute_Ks returns K(X*, X): rows = grid (GP_scripts.py:97-123)
        out["ref_Ks_loops_%d" % t] = (
            r * gs["compute_Ks"](X[:, 0], X[:, 1], X2[:, 0], X2[:, 1], ldf, 1)
            + (1 - r) * gs["compute_Ks"](X[:, 0], X[:, 1], X2[:, 0], X2[:, 1], lcf, 2))
        k.update_gradients_full(W_sym, X, None)
        out["ref_grad_compat_sym_%d" % t] = np.array(
            [k.length_df.gradient, k.length_cf.gradient, k.ratio.gradient])
        k.update_gradients_full(W_x, X, X2)
        out["ref_grad_compat_x_%d" % t] = np.array(
            [k.length_df.gradient, k.length_cf.gradient, k.ratio.gradient])
    # scalar squared-exponential helpers (GP_scripts.py:125-142; nonDivK's divFree=0 branch :67-68)
    out["ref_sqExp"] = gs["sqExp"](X[:, 0], X[:, 1], X2[:, 0], X2[:, 1], 1.7)
    out["ref_rbf_x"] = gs["rbf"](X[:, 0], X2[:, 0], l=1.3, sigma=0.8, noise=0.05)
    out["ref_rbf_sym"] = gs["rbf"](X[:, 0], X[:, 0], l=1.3, sigma=0.8, noise=0.05)
    nd = rs.nondivk_class(1.7)
    out["ref_nonDivK"] = nd.K(X, X2)
    out["ref_nonDivK_diag"] = nd.Kdiag(X2)
    nd.update_gradients_full(W_x, X, X2)
    out["ref_nonDivK_grad_compat"] = np.array([nd.length.gradient])
    nr = rs.nonrotk_class(0.9)
    out["ref_nonRotK"] = nr.K(X, X2)
    out["ref_nonRotK_diag"] = nr.Kdiag(X2)
    nr.update_gradients_full(W_x, X, X2)
    out["ref_nonRotK_grad_compat"] = np.array([nr.length_cf.gradient])
    np.savez_compressed(os.path.join(HERE, "kernel_small.npz"), **out)
    print("kernel_small.npz written")


def simlaser(ts, loops):
    """GP_laser.simLaser(ts) with the defaults l_df=l_cf=2, rate=.5, noise=.05
    (GP_laser.py:145-187), run through the reference's own functions."""
    tr = rs.load_simul_tracks()
    X, y, Xs = orc.simlaser_inputs(tr, ts=ts)          # host prep; geopy absent -> haversine
    l_df = l_cf = 2.0
    rate, noise = 0.5, 0.05
    gs = rs.gp_scripts()
    xo, yo = X[:, 0], X[:, 1]
    t0 = time.time()
    if loops:     # the literal loops of GP_laser.py:177,181 (minutes)
        K = rate * gs["compute_K"](xo, yo, l_df, 1) + (1 - rate) * gs["compute_K"](xo, yo, l_cf, 2)
        Ks_df = gs["compute_Ks"](xo, yo, Xs[:, 0], Xs[:, 1], l_df, 1)
        Ks_cf = gs["compute_Ks"](xo, yo, Xs[:, 0], Xs[:, 1], l_cf, 2)
    else:         # the vectorised reference function (GP_scripts.py:6-42), same values
        K = gs["myKernel"](X, X, l_df, l_cf, rate)
        Ks_df = gs["myKernel"](Xs, X, l_df, l_cf, 1.0)
        Ks_cf = gs["myKernel"](Xs, X, l_df, l_cf, 0.0)
    print("  reference kernels built in %.1f s" % (time.time() - t0))
    K = K + np.identity(np.size(K, 0)) * noise                      # GP_laser.py:178-179
    Ki = np.linalg.inv(K)                                           # GP_laser.py:180
    obs = np.reshape(y, [y.size, 1])
    Ks_simlaser = rate * Ks_df + (1 - rate) * rate * Ks_cf          # GP_laser.py:181 (as written)
    Ks_correct = rate * Ks_df + (1 - rate) * Ks_cf                  # GP_laser.py:122
    f_simlaser = gs["getMean"](Ks_simlaser, Ki, obs)                # GP_laser.py:183
    f_correct = gs["getMean"](Ks_correct, Ki, obs)
    Kss_diag = rate / l_df ** 2 + (1 - rate) / l_cf ** 2            # diag of GP_laser.py:128
    var = Kss_diag - np.einsum("ij,ij->i", Ks_correct @ Ki, Ks_correct)   # diag of GP_laser.py:129
    L = sla.cholesky(K, lower=True)
    alpha = sla.cho_solve((L, True), y)
    lml = -0.5 * y @ alpha - np.sum(np.log(np.diag(L))) - 0.5 * y.size * np.log(2 * np.pi)
    np.savez_compressed(
        os.path.join(HERE, "simlaser_ts%d.npz" % ts),
        X=X, y=y, Xs=Xs, theta=np.array([l_df, l_cf, rate]), noise=noise,
        ref_mean_simlaser=f_simlaser, ref_mean=f_correct, ref_var=var,
        derived_lml=lml, derived_cond=np.linalg.cond(K), used_loops=loops)
    print("simlaser_ts%d.npz written: lml=%.10f var in [%.5f, %.5f]" % (ts, lml, var.min(), var.max()))


def simlaser_tracks():
    """Raw track columns (lat, lon, u, v at ts = 0 and 100) so that GP_laser.simLaser can be
    driven end to end on the GPU box, where simulTracks.pkl is absent."""
    tr = rs.load_simul_tracks()
    cols = [0, 100]
    np.savez_compressed(os.path.join(HERE, "simlaser_tracks.npz"), ts=np.array(cols),
                        lat=np.asarray(tr.lat)[:, cols], lon=np.asarray(tr.lon)[:, cols],
                        u=np.asarray(tr.u)[:, cols], v=np.asarray(tr.v)[:, cols])


if __name__ == "__main__":
    assert rs.available(), "reference tree not found"
    kernel_small()
    simlaser(0, loops=True)
    simlaser(100, loops=False)
    simlaser_tracks()
