"""One fit + two predicts at configs[1] size (profiling target: ncu -k regex:predict).
    python tools/predict_once.py [mode] [N] [grid side]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp2d_b200 as gp                                   # noqa: E402
from gp2d_b200 import synthetic as syn                   # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
side = int(sys.argv[3]) if len(sys.argv) > 3 else 320
gp.set_predict_i8(mode)
if len(sys.argv) > 4:
    from gp2d_b200._lib import lib
    lib.gp2d_dbg_set_i8(int(sys.argv[4]))
X, y = syn.drifter_snapshot(N, config_id=2)
Xs = gp.as_dev(syn.prediction_grid(X, side, side))
m = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05)
m.fit()
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    mean, var = m.predict(Xs)
    b.record()
    torch.cuda.synchronize()
    print("predict %.3f ms  var[%g, %g]" % (a.elapsed_time(b), float(var.min()), float(var.max())))
