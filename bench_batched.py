"""configs[3] and configs[4] of BASELINE.json for bench.py (one measurement per invocation, added to the
bench line as `restarts` and `snapshots`).

restarts   hyper-parameter fit at N = 2000: log-marginal-likelihood + gradient, 64 restarts in total, restart r
           on rank r % world (strong scaling: 8 per GPU on 8 B200).  The restarts of a rank advance in lock
           step: every L-BFGS-B round is ONE gp2d_lml_grad_batched call (models.GPRegression.optimize_restarts,
           batched=True), dist.gather_best picks the winner (krig.py:450; GP_plots.py:765).
snapshots  time series of independent snapshots, N = 8192 each, 100 x 100 grid, 64 snapshots PER RANK (weak
           scaling: 512 on 8 B200, the configuration BASELINE.json names) through the package API
           gp2d_b200.krig_snapshots (numpy in, numpy out; fits batched two at a time, fused predict per
           snapshot; krig.py:541-557).
"""
from __future__ import annotations

import time

import numpy as np

RESTARTS_TOTAL = 64
RESTART_N = 2000
RESTART_MAXITER = 100
SNAP_PER_RANK = 64
SNAP_N = 8192
SNAP_GRID = (100, 100)
THETA = (1.3, 3.1, 0.2)
NOISE = 0.05


def run(torch, dist, gp, gdist, syn, rank, world, dev, peak_tf):
    from gp2d_b200 import models, myKernel

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def tmax(s):
        t = torch.tensor([s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def tsum(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return float(t.item())

    out = {}

    # ---- configs[3]: restarts ---------------------------------------------------------------------
    X, y = syn.drifter_snapshot(RESTART_N, config_id=4)
    model = models.GPRegression(X, y[:, None], myKernel.myKernel(2, [0, 1], 1.0, 1.0, 0.5), noise_var=0.1)
    g = model._gp
    g.lml_and_grad()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        g.lml_and_grad()
    t_single = (time.perf_counter() - t0) / 5
    mine = len([r for r in range(RESTARTS_TOTAL) if r % world == rank])
    # one batched evaluation of this rank's share, timed alone
    hb = gp.HelmholtzBatch(X, y, B=mine, jitter=model.jitter)
    th = np.tile(np.array([1.0, 1.0, 0.5, 0.1]), (mine, 1)) * (1.0 + 0.01 * np.arange(mine))[:, None]
    th[:, 2] = 0.5
    hb.lml_and_grad(th)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        hb.lml_and_grad(th)
    t_batch = (time.perf_counter() - t0) / 3
    del hb
    torch.cuda.empty_cache()
    sync()
    t0 = time.perf_counter()
    model.optimize_restarts(num_restarts=RESTARTS_TOTAL, verbose=False, seed=4, max_iters=RESTART_MAXITER,
                            rank=rank, world=world, batched=True)
    best = gdist.gather_best(model)
    sync()
    t_restarts = tmax(time.perf_counter() - t0)
    nfev = tsum(sum(r.funct_eval for r in model.optimization_runs))
    rounds = tmax(model.lockstep_stats["rounds"] if mine else 0)
    n = 2 * RESTART_N
    flop_eval = float(n) ** 3          # potrf n^3/3 + inverse n^3/3 + K^-1 = Z^T Z n^3/3
    out["restarts"] = {
        "workload": "configs[3]: LML + gradient at N=%d, %d restarts (L-BFGS-B, maxiter %d) sharded r %% world over %d GPU(s), "
                    "lock-step batches through gp2d_lml_grad_batched" % (RESTART_N, RESTARTS_TOTAL, RESTART_MAXITER, world),
        "scaling": "strong", "n_gpus": world, "restarts": RESTARTS_TOTAL, "restarts_per_gpu": mine,
        "wall_s": t_restarts, "restarts_per_s": RESTARTS_TOTAL / t_restarts,
        "objective_evaluations": int(nfev), "lockstep_rounds_max": int(rounds),
        "ms_per_evaluation_amortised": t_restarts / max(nfev, 1) * 1e3 * world,
        "single_evaluation_ms": t_single * 1e3,
        "batched_evaluation_ms": t_batch * 1e3, "batched_evaluation_ms_per_problem": t_batch * 1e3 / max(mine, 1),
        "batched_TFLOPps": mine * flop_eval / t_batch / 1e12, "batched_frac_of_fp64_peak": mine * flop_eval / t_batch / 1e12 / peak_tf,
        "best_objective": best, "theta": model.param_array.tolist(),
    }
    del model, g
    torch.cuda.empty_cache()

    # ---- configs[4]: snapshots --------------------------------------------------------------------
    S = SNAP_PER_RANK
    first = rank * S
    Xb = np.empty((S, SNAP_N, 2))
    yb = np.empty((S, 2 * SNAP_N))
    for s in range(S):
        Xb[s], yb[s] = syn.drifter_snapshot(SNAP_N, config_id=5, seed_offset=first + s)
    grid = syn.prediction_grid(Xb[0], *SNAP_GRID)
    gp.krig_snapshots(Xb[:2], yb[:2], grid, *THETA, NOISE, batch=2)         # warm-up
    sync()
    t0 = time.perf_counter()
    mean, var, lml = gp.krig_snapshots(Xb, yb, grid, *THETA, NOISE, batch=2)
    torch.cuda.synchronize()
    t_snap = tmax(time.perf_counter() - t0)
    n = 2 * SNAP_N
    M = grid.shape[0]
    flop = 2.0 * float(n) ** 3 / 3 + float(n) * n * 2 * M
    total = S * world
    out["snapshots"] = {
        "workload": "configs[4]: independent snapshots of N=%d (n=%d), %dx%d grid, %d per GPU on %d GPU(s) = %d "
                    "(512 on 8 B200); gp2d_b200.krig_snapshots, numpy in / out" % (SNAP_N, n, SNAP_GRID[0], SNAP_GRID[1], S, world, total),
        "scaling": "weak", "n_gpus": world, "snapshots": total, "wall_s": t_snap,
        "s_per_snapshot": t_snap / total, "snapshots_per_s": total / t_snap,
        "TFLOPps_aggregate": total * flop / t_snap / 1e12, "TFLOPps_note": "fp64-equivalent (fit 2n^3/3 on the FP64 pipe + predict n^2 m on the int8-sliced kernel)",
        "x_fp64_pipe_peak": total * flop / t_snap / 1e12 / (world * peak_tf),
        "h2d_bytes_per_snapshot": 8 * (2 * SNAP_N + 2 * SNAP_N), "d2h_bytes_per_snapshot": 8 * (4 * M + 1),
        "lml_first": float(lml[0]), "var_min": float(var.min()), "var_max": float(var.max()),
    }
    return out
