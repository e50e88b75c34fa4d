"""Oracle numbers at BASELINE.json's Target size (configs[2]: N = 16 384 observations, n = 32 768), one-off.

    python tools/config3_parity.py [N] [points] > profiles/r02_config3_parity.json

The oracle (numpy / scipy restatement of GP_laser.py:113-140 and myKernel.py:27-53) assembles the
32 768 x 32 768 covariance in chunks (8.6 GB), factorises it with LAPACK dpotrf on the host cores and
predicts at `points` sampled points of the 1000 x 1000 grid; the CUDA path fits the same snapshot and
predicts the same points.  Tolerances of BASELINE.json: LML 1e-6, mean / variance 1e-8 relative."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp2d_b200 as gp
from gp2d_b200 import synthetic
from oracle import gp_oracle as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P = int(sys.argv[2]) if len(sys.argv) > 2 else 200
theta, noise = (1.3, 3.1, 0.2), 0.05
X, y = synthetic.drifter_snapshot(N, config_id=3)
grid = synthetic.prediction_grid(X, 1000, 1000)
pick = np.random.default_rng(3).choice(grid.shape[0], P, replace=False)
pts = grid[pick]

m = gp.HelmholtzGP(X, y, *theta, noise)
t0 = time.perf_counter()
lml = m.fit()
mean, var = m.predict(pts)
torch.cuda.synchronize()
t_gpu = time.perf_counter() - t0
mean, var = mean.cpu().numpy(), var.cpu().numpy()

t0 = time.perf_counter()
f = orc.fit_chunked(X, y, *theta, noise, chunk=512)
t_fit = time.perf_counter() - t0
mo, vo = orc.predict(X, f, *theta, pts, chunk=100)
t_cpu = time.perf_counter() - t0

rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
out = {
    "what": "configs[2] parity: CUDA path vs oracle at N=%d (n=%d), %d sampled points of the 1000x1000 grid" % (N, 2 * N, P),
    "lml_gpu": lml, "lml_oracle": f["lml"], "lml_rel_err": abs(lml - f["lml"]) / abs(f["lml"]), "lml_tol": 1e-6,
    "mean_max_abs_err_over_max_abs": float(np.abs(mean - mo).max() / np.abs(mo).max()),
    "mean_rel_err_max": rel(mean, mo), "var_rel_err_max": rel(var, vo), "mean_var_tol": 1e-8,
    "var_min": float(vo.min()), "var_max": float(vo.max()),
    "gpu_fit_predict_s": t_gpu, "oracle_fit_s": t_fit, "oracle_total_s": t_cpu, "host_cpus": os.cpu_count(),
    "pass": bool(abs(lml - f["lml"]) <= 1e-6 * abs(f["lml"]) and
                 np.allclose(mean, mo, rtol=1e-8, atol=1e-9 * np.abs(mo).max()) and np.allclose(var, vo, rtol=1e-8, atol=1e-12)),
}
print(json.dumps(out))
