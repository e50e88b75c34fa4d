"""Bring-up check of the int8-sliced predictive kernel against the fp64 kernel and the oracle.
    python tools/i8_check.py [N] [grid side]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp2d_b200 as gp                                   # noqa: E402
from gp2d_b200 import synthetic as syn                   # noqa: E402
from oracle import gp_oracle as orc                      # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
side = int(sys.argv[2]) if len(sys.argv) > 2 else 320
theta, noise = (1.3, 3.1, 0.2), 0.05
X, y = syn.drifter_snapshot(N, config_id=2)
Xs = syn.prediction_grid(X, side, side)
M = Xs.shape[0]
res = {}
for mode in (1, 6, 7, 0):
    gp.set_predict_i8(mode)
    m = gp.HelmholtzGP(X, y, *theta, noise)
    Xsd = gp.as_dev(Xs)
    m.fit()
    mean, var = m.predict(Xsd)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mean, var = m.predict(Xsd)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t0 = time.perf_counter(); m.fit_async(); torch.cuda.synchronize(); tf = time.perf_counter() - t0
    res[mode] = (mean.cpu().numpy(), var.cpu().numpy())
    n = 2 * N
    fl = float(n) * n * 2 * M
    print("mode %d: predict %.2f ms (%.1f TFLOP/s fp64-equivalent), fit %.2f ms, var in [%.3e, %.3e]" %
          (mode, min(ts), fl / (min(ts) * 1e-3) / 1e12, tf * 1e3, res[mode][1].min(), res[mode][1].max()), flush=True)
m1, v1 = res[1]
for mode in (6, 7, 0):
    mm, vv = res[mode]
    print("mode %d vs fp64 kernel: mean max abs diff %.3e (rel to max|mean| %.3e), var max rel diff %.3e" %
          (mode, np.abs(mm - m1).max(), np.abs(mm - m1).max() / np.abs(m1).max(), np.abs(vv / v1 - 1).max()))
# oracle on a sample
idx = np.random.default_rng(0).choice(M, 400, replace=False)
f = orc.fit(X, y, *theta, noise)
mo, vo = orc.predict(X, f, *theta, Xs[idx])
sel = np.concatenate([idx, M + idx])
for mode in (1, 6, 7):
    mm, vv = res[mode]
    print("mode %d vs oracle (400 points): mean rel %.3e, var rel %.3e" %
          (mode, np.abs(mm[sel] - mo).max() / np.abs(mo).max(), np.abs(vv[sel] / vo - 1).max()))
# partition invariance of the int8 path
gp.set_predict_i8(6)
m = gp.HelmholtzGP(X, y, *theta, noise)
m.fit()
mA, vA = m.predict(gp.as_dev(Xs[:1000]))
mB, vB = m.predict(gp.as_dev(Xs[37:537]))
print("partition invariance (bitwise):", bool(torch.equal(mA[37:537], mB[:500]) and torch.equal(vA[37:537], vB[:500])
                                              and torch.equal(mA[1000 + 37:1000 + 537], mB[500:])))
gp.set_predict_i8(0)
