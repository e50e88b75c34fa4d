"""Zero-slice skipping of the int8 predictive kernel: bit-identical to the dense schedule, and how much faster.
    python tools/i8_skip_check.py [N] [grid side]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp2d_b200 as gp                                   # noqa: E402
from gp2d_b200 import synthetic as syn                   # noqa: E402
from gp2d_b200._lib import lib                           # noqa: E402

lib.gp2d_dbg_set_i8.restype = C.c_int
lib.gp2d_dbg_set_i8.argtypes = [C.c_int]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
side = int(sys.argv[2]) if len(sys.argv) > 2 else 320
theta, noise = (1.3, 3.1, 0.2), 0.05
X, y = syn.drifter_snapshot(N, config_id=2)
Xs = syn.prediction_grid(X, side, side)
M = Xs.shape[0]
Xsd = gp.as_dev(Xs)
out = {}
for s in (6, 7):
    gp.set_predict_i8(s)
    m = gp.HelmholtzGP(X, y, *theta, noise)
    m.fit()
    for dbg in (8, 0):
        lib.gp2d_dbg_set_i8(dbg)
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
        out[(s, dbg)] = (mean.clone(), var.clone())
        print("S=%d %s: predict %.2f ms" % (s, "dense schedule" if dbg else "zero slices skipped", min(ts)), flush=True)
    lib.gp2d_dbg_set_i8(0)
    same = torch.equal(out[(s, 0)][0], out[(s, 8)][0]) and torch.equal(out[(s, 0)][1], out[(s, 8)][1])
    print("S=%d skipped == dense bit for bit: %s" % (s, same), flush=True)
gp.set_predict_i8(1)
m = gp.HelmholtzGP(X, y, *theta, noise)
m.fit()
m1, v1 = m.predict(Xsd)
for s in (6, 7):
    mm, vv = out[(s, 0)]
    print("S=%d vs fp64 kernel: mean rel %.3e, var max rel %.3e" % (s, float((mm - m1).abs().max() / m1.abs().max()), float((vv / v1 - 1).abs().max())))
gp.set_predict_i8(0)
