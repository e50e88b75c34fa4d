"""One fit (+ optional lml_grad) at N obs, for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gp2d_b200 as gp
from gp2d_b200 import synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
X, y = synthetic.drifter_snapshot(N, config_id=2)
m = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05)
for _ in range(reps):
    m.fit_async()
torch.cuda.synchronize()
print("lml", m.fit())
