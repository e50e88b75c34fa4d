"""BASELINE.json configs[2] on one GPU: N=16384 observations (32768 x 32768 fp64 covariance).

Times the stages the north-star targets name -- covariance build GB/s, Cholesky TFLOP/s, the
full fit (Cholesky + L^-1 + alpha + LML), LML+gradient -- and the fused predictive pass on a
SHARD of the 1M-point grid (predict cost is exactly linear in grid points), and checks
size-independent properties.  Usage: python tools/config3.py [N] [grid_shard_points]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp2d_b200 as gp
from gp2d_b200 import synthetic
from gp2d_b200._lib import lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
MSH = int(sys.argv[2]) if len(sys.argv) > 2 else 64 * 148
dev = torch.device("cuda:0")
theta, noise = (1.3, 3.1, 0.2), 0.05
n = 2 * N


def ev():
    return torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=2, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


out = {"N": N, "n": n}
X, y = synthetic.drifter_snapshot(N, config_id=3)
side = int(np.ceil(np.sqrt(1000000)))
Xs_full = synthetic.prediction_grid(X, side, side)
Xd, yd = gp.as_dev(X), gp.as_dev(y)
st = torch.cuda.current_stream().cuda_stream

# covariance build, reference block layout, full matrix (8 n^2 bytes written)
K = torch.empty((n, n), dtype=torch.float64, device=dev)
t = timeit(lambda: gp.kernel_K(Xd, None, *theta, diag_add=noise, out=K))
out["build_block_ms"] = t * 1e3
out["build_block_GBps"] = 8.0 * n * n / t / 1e9

# Cholesky alone (public potrf on the block-layout matrix; in place -> rebuild each time)
nb_ws = lib.gp2d_potrf_workspace_bytes(n)
ws = torch.empty(nb_ws, dtype=torch.uint8, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)


def potrf():
    gp.kernel_K(Xd, None, *theta, diag_add=noise, out=K)
    lib.gp2d_potrf(K.data_ptr(), n, n, ws.data_ptr(), nb_ws, info.data_ptr(), st)


t_p = timeit(potrf, reps=2) - t
out["potrf_ms"] = t_p * 1e3
out["potrf_TFLOPps"] = n ** 3 / 3.0 / t_p / 1e12
out["potrf_info"] = int(info.item())
# residual check on a sample of entries: (L L^T)[i,j] == K[i,j]
L = torch.tril(K)
idx = torch.randint(0, n, (64,), device=dev)
Kchk = gp.kernel_K(Xd, None, *theta, diag_add=noise)
R = L[idx] @ L.t() - Kchk[idx]
out["potrf_max_abs_residual_rows"] = float(R.abs().max())
del L, Kchk, R, K, ws
torch.cuda.empty_cache()

m = gp.HelmholtzGP(Xd, yd, *theta, noise)
t_fit = timeit(lambda: m.fit_async(), reps=2)
out["fit_ms"] = t_fit * 1e3
out["fit_TFLOPps_2n3_3"] = 2.0 * n ** 3 / 3.0 / t_fit / 1e12
lml = m.fit()
out["lml"] = lml
Xs = gp.as_dev(Xs_full[:MSH])
t_pr = timeit(lambda: m.predict(Xs), reps=1)
fl = float(n) * n * 2 * MSH + 2.0 * n * 2 * MSH
out["predict_shard_points"] = MSH
out["predict_shard_ms"] = t_pr * 1e3
out["predict_TFLOPps"] = fl / t_pr / 1e12
out["predict_1M_grid_extrapolated_s"] = t_pr * 1000000 / MSH
# GP identity at a sample of observation sites: K alpha = y - noise alpha
sel = np.random.default_rng(0).choice(N, 256, replace=False)
pm, pv = m.predict(X[sel])
al = m.alpha()
rhs = torch.cat([m.y[sel], m.y[N + sel]]) - noise * torch.cat([al[sel], al[N + sel]])
out["identity_max_abs_err"] = float((pm - rhs).abs().max())
out["var_min"], out["var_max"] = float(pv.min()), float(pv.max())
t_g = timeit(lambda: m.lml_and_grad(), reps=1, warm=0)
out["lml_grad_ms"] = t_g * 1e3
print(json.dumps(out))
