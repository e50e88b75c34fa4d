"""BASELINE.json configs[3] and configs[4], sharded over the ranks of a torchrun launch.

configs[3]: hyper-parameter fit -- log-marginal-likelihood + gradient, restarts sharded
            round-robin over the ranks (models.GPRegression.optimize_restarts + dist.gather_best).
configs[4]: time series of independent snapshots (N=8192 each, 100x100 grid) batch-kriged,
            snapshots sharded round-robin.

    python tools/configs45.py [--restarts R] [--n-fit N] [--snapshots S] [--n-snap N] [--maxiter I]
    python -m torch.distributed.run --nproc-per-node G ... tools/configs45.py ...

Bounded defaults (a few restarts / snapshots per rank); per-unit times are what scale."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp2d_b200 as gp
from gp2d_b200 import dist as gdist
from gp2d_b200 import models, myKernel, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--restarts", type=int, default=8)
ap.add_argument("--n-fit", type=int, default=2000)
ap.add_argument("--maxiter", type=int, default=100)
ap.add_argument("--parallel", type=int, default=4, help="concurrent restarts per GPU (own stream + workspace each)")
ap.add_argument("--snapshots", type=int, default=8)
ap.add_argument("--n-snap", type=int, default=8192)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def tmax(s):
    t = torch.tensor([s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"n_gpus": world}

# ---- configs[3]: restarts -------------------------------------------------------------------
X, y = synthetic.drifter_snapshot(args.n_fit, config_id=4)
model = models.GPRegression(X, y[:, None], myKernel.myKernel(2, [0, 1], 1.0, 1.0, 0.5), noise_var=0.1)
g = model._gp
g.lml_and_grad()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    g.lml_and_grad()
t_eval = (time.perf_counter() - t0) / 5
sync()
t0 = time.perf_counter()
model.optimize_restarts(num_restarts=args.restarts, verbose=False, seed=4, max_iters=args.maxiter, rank=rank, world=world,
                        parallel=args.parallel)
best = gdist.gather_best(model)
sync()
t_restarts = tmax(time.perf_counter() - t0)
nfev = sum(r.funct_eval for r in model.optimization_runs)
nf = torch.tensor([nfev], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(nf)
out["restarts"] = {
    "N": args.n_fit, "restarts": args.restarts, "maxiter": args.maxiter, "parallel_per_gpu": args.parallel, "lml_grad_eval_ms": t_eval * 1e3,
    "wall_s": t_restarts, "restarts_per_s": args.restarts / t_restarts, "objective_evaluations": int(nf.item()),
    "best_objective": best, "theta": model.param_array.tolist()}

# ---- configs[4]: batched snapshots ------------------------------------------------------------
mine = gdist.round_robin(args.snapshots, rank, world)
N = args.n_snap
snaps = []
for sidx in mine:
    Xs_, ys_ = synthetic.drifter_snapshot(N, config_id=5, seed_offset=sidx)
    snaps.append((gp.as_dev(Xs_), gp.as_dev(ys_), gp.as_dev(synthetic.prediction_grid(Xs_, 100, 100))))
m = gp.HelmholtzGP(snaps[0][0], snaps[0][1], 1.3, 3.1, 0.2, 0.05) if snaps else None
mean = torch.empty(20000, dtype=torch.float64, device=dev)
var = torch.empty(20000, dtype=torch.float64, device=dev)
if m is not None:                       # warm-up
    m.fit_async(); m.predict(snaps[0][2], out_mean=mean, out_var=var)
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lmls = []
for (Xd, yd, Gd) in snaps:
    m.X, m.y = Xd, yd
    m.fit_async()
    m.predict(Gd, out_mean=mean, out_var=var)
    lmls.append(m._scal[:1].clone())
e1.record()
sync()
t_snap = tmax(e0.elapsed_time(e1) * 1e-3)
n = 2 * N
out["snapshots"] = {
    "N": N, "grid_points": 10000, "snapshots": args.snapshots, "device_s": t_snap,
    "s_per_snapshot": t_snap / max(1, args.snapshots), "snapshots_per_s": args.snapshots / t_snap,
    "flop_per_snapshot": 2.0 * n ** 3 / 3 + float(n) * n * 20000,
    "TFLOPps_aggregate": args.snapshots * (2.0 * n ** 3 / 3 + float(n) * n * 20000) / t_snap / 1e12,
    "lml_first": float(lmls[0].item()) if lmls else None, "info": int(m._info.item()) if m is not None else None}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
