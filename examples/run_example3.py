#!/usr/bin/env python
"""The reference's end-to-end example (GP_plots.run_example3, GP_plots.py:705-779) on the GPU path:
sample a synthetic divergence-free (or curl-free) velocity field at random points, reconstruct it
(a) component by component with a product of two 1-D RBF kernels and (b) jointly with the
Helmholtz kernel ``myKernel(2, [0, 1], 0.6, 0.6, 1)`` after ``optimize_restarts``, and compare the
RMSE of both reconstructions on the full grid.

    python examples/run_example3.py [nsamples] [divFree] [num_restarts]

The reference runs 1500 restarts (its analytic length-scale gradient is wrong);
with the correct gradient a handful is enough."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gp2d_b200 import GP_scripts, kern, models              # noqa: E402
from gp2d_b200.myKernel import myKernel                     # noqa: E402


def run_example3(nsamples=5, divFree=1, rms=0, num_restarts=8, seed=0):
    x, y, phi, xm, ym, um, vm = GP_scripts.generate_2D_gaussian(divFree)
    X1, X2 = np.meshgrid(xm, ym)
    X1 = np.reshape(X1, [X1.size, 1])
    X2 = np.reshape(X2, [X2.size, 1])
    um2, vm2 = np.reshape(um, [-1]), np.reshape(vm, [-1])
    samples = np.random.default_rng(seed).integers(0, X1.size, nsamples)
    x1, x2 = X1[samples], X2[samples]
    y1, y2 = um2[samples][:, None], vm2[samples][:, None]
    X = np.concatenate([x1, x2], axis=1)
    Y = np.concatenate([y1, y2], axis=0)
    Xg = np.concatenate([X1, X2], axis=1)
    # component by component: RBF(x) * RBF(y)  (GP_plots.py:738-756)
    k = kern.RBF(1, active_dims=[0], variance=5, lengthscale=0.1) * kern.RBF(1, active_dims=[1], variance=5, lengthscale=0.1)
    model_u = models.GPRegression(X, y1, k.copy())
    model_u.optimize()
    model_u.optimize_restarts(num_restarts=num_restarts, verbose=False, seed=seed)
    ug = np.reshape(model_u.predict(Xg)[0], [xm.size, -1])
    model_v = models.GPRegression(X, y2, k.copy())
    model_v.optimize()
    model_v.optimize_restarts(num_restarts=num_restarts, verbose=False, seed=seed)
    vg = np.reshape(model_v.predict(Xg)[0], [xm.size, -1])
    # jointly, with the divergence-free + curl-free kernel  (GP_plots.py:759-770)
    model = models.GPRegression(X, Y, myKernel(2, [0, 1], 0.6, 0.6, 1))
    model.optimize_restarts(num_restarts=num_restarts, verbose=False, seed=seed)
    f, fVar = model.predict(Xg)
    fu = np.reshape(f[:f.size // 2], [xm.size, -1])
    fv = np.reshape(f[f.size // 2:], [xm.size, -1])
    rmsug, rmsvg = GP_scripts.rmse(X1, X2, ug, vg, X1, X2, um, vm, knd='')
    rmsfu, rmsfv = GP_scripts.rmse(X1, X2, fu, fv, X1, X2, um, vm, knd='')
    if rms == 0:
        return xm, ym, um, vm, ug, vg, fu, fv, rmsug, rmsvg, rmsfu, rmsfv, model_u.param_array, model_v.param_array, model.param_array
    return rmsug, rmsvg, rmsfu, rmsfv, model_u.param_array, model_v.param_array, model.param_array


if __name__ == "__main__":
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    df = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    nr = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    out = run_example3(ns, df, rms=1, num_restarts=nr)
    print("RMSE scalar RBF (u, v): %.5f %.5f   Helmholtz kernel (u, v): %.5f %.5f" % out[:4])
    print("hyper-parameters [length_df, length_cf, ratio, noise]:", out[6])
