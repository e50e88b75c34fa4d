"""Accuracy of the fused path against the oracle as the conditioning bound n k**/noise grows, with the
plain and the robust (refined-panel) factorisation.  Usage: python tests/tools/cond_sweep.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gp2d_b200 as gp
from gp2d_b200._lib import lib
from oracle import gp_oracle as orc

lib.gp2d_dbg_set_robust_cond.argtypes = [C.c_double]
rng = np.random.default_rng(0)
N, M = 600, 800
X = rng.uniform(0, 12, (N, 2))
Xs = rng.uniform(0, 12, (M, 2))
f_u = lambda P: np.sin(P[:, 0] / 2) * np.cos(P[:, 1] / 3)
f_v = lambda P: np.cos(P[:, 0] / 3)
th = (1.3, 3.1, 0.2)
print("%8s %9s | %-32s | %-32s | %s" % ("noise", "bound", "plain: mean / var / lml", "robust: mean / var / lml", "refined predict: mean / var"))
for noise in (1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8):
    y = np.concatenate([f_u(X), f_v(X)]) + np.sqrt(noise) * rng.normal(size=2 * N)
    fo = orc.fit(X, y, *th, noise)
    mo, vo = orc.predict(X, fo, *th, Xs)
    row = []
    for thr in (1e300, 0.0):
        lib.gp2d_dbg_set_robust_cond(thr)
        g = gp.HelmholtzGP(X, y, *th, noise)
        lml = g.fit()
        m, v = g.predict(Xs)
        row.append("%.1e / %.1e / %.1e" % (np.abs(m.cpu().numpy() - mo).max() / np.abs(mo).max(),
                                           np.abs(v.cpu().numpy() - vo).max() / vo.max(), abs(lml - fo["lml"]) / abs(fo["lml"])))
    mr, vr = g.predict_refined(Xs)
    ref = "%.1e / %.1e" % (np.abs(mr.cpu().numpy() - mo).max() / np.abs(mo).max(), np.abs(vr.cpu().numpy() - vo).max() / vo.max())
    print("%8.0e %9.1e | %-32s | %-32s | %s" % (noise, g.cond_bound(), row[0], row[1], ref))
lib.gp2d_dbg_set_robust_cond(1e7)
