"""Steady-state time of one fit (gp2d_fit: spatial order, build, Cholesky + inverse, pack, digit slices, alpha, LML).
    python tools/fit_time.py [N] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp2d_b200 as gp                                   # noqa: E402
from gp2d_b200 import synthetic                          # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
X, y = synthetic.drifter_snapshot(N, config_id=2)
m = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05)
for _ in range(5):
    m.fit_async()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    m.fit_async()
b.record()
torch.cuda.synchronize()
print("N=%d: fit %.3f ms (mean of %d), lml %.12g" % (N, a.elapsed_time(b) / reps, reps, m.fit()))
