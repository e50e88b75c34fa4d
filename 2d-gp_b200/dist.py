"""Sharding of the three independent axes of the path across ranks: grid
points, hyper-parameter restarts, snapshots.  One process per GPU; torch.distributed
(NCCL over NVLink on GPUs, gloo in the CPU tests) is used only to gather results -- there is
no collective inside the data path, a single factorisation stays on one GPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .engine import ROBUST_COND


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world_size, align=64):
    """Contiguous [lo, hi) slice of n items for ``rank``; boundaries are multiples of
    ``align`` (64 grid points = one predict column tile) so every rank runs whole tiles."""
    units = -(-n // align)
    per, extra = divmod(units, world_size)
    lo_u = rank * per + min(rank, extra)
    hi_u = lo_u + per + (1 if rank < extra else 0)
    return min(lo_u * align, n), min(hi_u * align, n)


def shard_cyclic(n, rank, world_size, block=320):
    """Block-cyclic share of n grid points for ``rank``: blocks of ``block`` points (whole column tiles of both
    predictive kernels: 40 and 64 points) dealt round robin.  The cost of a column tile is far from uniform -- the
    product Z K* is lower triangular, so a tile whose covariances concentrate on early observations meets a non-zero
    K* slice in every row block below them, one at the far end in the last few only (with the zero-slice skip of
    predict_i8.cu contiguous halves of the configs[2] grid split the work 2 : 1) -- and interleaving evens it out."""
    blocks = np.arange(rank, -(-n // block), world_size, dtype=np.int64)
    idx = (blocks[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < n]


def round_robin(n, rank, world_size):
    """Indices rank, rank+world, ... : restarts and snapshots."""
    return list(range(rank, n, world_size))


def gather_concat(local: torch.Tensor, sizes=None):
    """all-gather variable-length 1-D tensors and concatenate in rank order."""
    rank, ws = world()
    if ws == 1:
        return local
    n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(ws)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    buf = torch.zeros(mx, dtype=local.dtype, device=local.device)
    buf[:local.numel()] = local
    parts = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


def broadcast_fit(gp, src=0, chunk_bytes=1 << 30):
    """Ship the fit state of rank ``src`` (L^-1 tiles, alpha, X: HelmholtzGP.predict_state) to
    every rank, so that one factorisation serves grid shards on all GPUs (the factorisation
    itself is never distributed).  Broadcast in <= 1 GiB pieces."""
    rank, ws = world()
    if ws == 1:
        return
    state = gp.predict_state()
    for lo in range(0, state.numel(), chunk_bytes):
        dist.broadcast(state[lo:lo + chunk_bytes], src=src)
    gp.fitted = True


def predict_sharded(gp, Xs, include_noise=False, refined=None):
    """Grid-sharded prediction: rank r predicts its block-cyclic share of the grid (shard_cyclic) with the
    fit state it holds (every rank fits the same snapshot, or receives it by broadcast), then
    mean/var shards are all-gathered and put back in grid order.  Returns (mean[2M], var[2M]) on every rank
    (the same bits as one rank predicting the whole grid: the kernels are partition invariant).  ``refined``:
    None takes the iterated solve for ill-conditioned covariances (engine.refined_predict), like
    GPRegression.predict; True / False force it."""
    rank, ws = world()
    Xs = np.asarray(Xs, dtype=np.float64) if not isinstance(Xs, torch.Tensor) else Xs
    M = Xs.shape[0]
    if refined is None:
        refined = hasattr(gp, "cond_bound") and gp.cond_bound() > ROBUST_COND
    fn = gp.predict_refined if refined else gp.predict
    if ws == 1:
        return fn(Xs, include_noise=include_noise)
    idx = shard_cyclic(M, rank, ws)
    sel = torch.as_tensor(idx, device=Xs.device) if isinstance(Xs, torch.Tensor) else idx
    mean, var = fn(Xs[sel], include_noise=include_noise)
    m = idx.shape[0]
    # shards arrive concatenated in rank order: scatter them back to grid order
    order = torch.as_tensor(np.concatenate([shard_cyclic(M, r, ws) for r in range(ws)]), device=mean.device)
    out = []
    for t in (mean, var):
        full = torch.empty(2 * M, dtype=t.dtype, device=t.device)
        full[order] = gather_concat(t[:m].contiguous())
        full[M + order] = gather_concat(t[m:].contiguous())
        out.append(full)
    return out[0], out[1]


def gather_best(model):
    """After model.optimize_restarts(..., rank, world): pick the globally best run and load
    its parameters on every rank.  Exchanges (f_opt, x_opt) only."""
    rank, ws = world()
    runs = model.optimization_runs
    nfree = len(model._free_params())
    if runs:
        best = min(runs, key=lambda o: o.f_opt)
        rec = np.concatenate([[best.f_opt], best.x_opt])
    else:
        rec = np.concatenate([[np.inf], np.zeros(nfree)])
    if ws > 1:
        dev = model._gp.device if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(rec, dtype=torch.float64, device=dev)
        allr = [torch.empty_like(t) for _ in range(ws)]
        dist.all_gather(allr, t)
        recs = torch.stack(allr).cpu().numpy()
        rec = recs[np.argmin(recs[:, 0])]
    if not np.isfinite(rec[0]):
        # no rank has a successful run (robust=True swallows failing restarts): keep the model as it is
        import warnings
        warnings.warn("gather_best: no successful optimisation run on any rank; parameters left unchanged", RuntimeWarning)
        return float("inf")
    model._set_free(rec[1:])
    model.parameters_changed()
    return float(rec[0])
