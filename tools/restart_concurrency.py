"""Throughput of hyper-parameter restarts vs the number run concurrently on one GPU (configs[3]):
each worker has its own CUDA stream and fit workspace (models.GPRegression.optimize_restarts(parallel=)).
One evaluation is n^3 flop (1.8 ms at DMMA peak for N=2000) inside a 6 ms chain of mostly small kernels."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp2d_b200 import models, myKernel, synthetic
X, y = synthetic.drifter_snapshot(2000, config_id=4)
for par in (1, 2, 4, 8):
    m = models.GPRegression(X, y[:, None], myKernel.myKernel(2, [0, 1], 1.0, 1.0, 0.5), noise_var=0.1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.optimize_restarts(num_restarts=16, verbose=False, seed=4, max_iters=100, parallel=par)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    nf = sum(r.funct_eval for r in m.optimization_runs)
    print("parallel=%d: %.3f s for 16 restarts, %d evaluations, %.2f ms per evaluation (wall/evals)" % (par, dt, nf, dt / nf * 1e3))
