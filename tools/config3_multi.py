"""BASELINE.json configs[2] across GPUs: ONE factorisation (rank 0), its predict state broadcast
over NCCL, the 1M-point grid sharded over the ranks, mean/variance all-gathered.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29517 tools/config3_multi.py [N] [grid_points]

Device times are CUDA-event times, max over ranks.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp2d_b200 as gp
from gp2d_b200 import dist as gdist
from gp2d_b200 import synthetic

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
MG = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
theta, noise = (1.3, 3.1, 0.2), 0.05
X, y = synthetic.drifter_snapshot(N, config_id=3)
side = int(np.ceil(np.sqrt(MG)))
Xs = synthetic.prediction_grid(X, side, side)[:MG]


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def tmax(ms):
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


m = gp.HelmholtzGP(X, y, *theta, noise)
lo, hi = gdist.shard_range(MG, rank, world)
Xsd = gp.as_dev(Xs[lo:hi])
m.predict(Xsd[:64])                  # warm-up (module load, scratch allocation); state is rank-local junk
torch.cuda.synchronize()
if world > 1:                        # NCCL connection set-up outside the timed region
    dist.broadcast(torch.zeros(1 << 20, dtype=torch.uint8, device=dev), src=0)
    dist.barrier()
e0 = ev()
if rank == 0:
    m.fit_async()
e1 = ev()
gdist.broadcast_fit(m, src=0)
e2 = ev()
mean, var = m.predict(Xsd)
e3 = ev()
mu = [gdist.gather_concat(mean[:hi - lo].contiguous()), gdist.gather_concat(mean[hi - lo:].contiguous())]
vv = [gdist.gather_concat(var[:hi - lo].contiguous()), gdist.gather_concat(var[hi - lo:].contiguous())]
e4 = ev()
torch.cuda.synchronize()
# the receiving ranks sit in the broadcast while rank 0 is still factorising: the transfer itself is
# rank 0's own broadcast time
t_fit, t_pred, t_g = tmax(e0.elapsed_time(e1)), tmax(e2.elapsed_time(e3)), tmax(e3.elapsed_time(e4))
t_bc = tmax(e1.elapsed_time(e2) if rank == 0 else 0.0)
total = tmax(e0.elapsed_time(e4))
n = 2 * N
if rank == 0:
    state_bytes = m.predict_state().numel()
    print(json.dumps({
        "config": "configs[2]: N=%d obs (n=%d), %d-point grid sharded over %d GPU(s), one factorisation" % (N, n, MG, world),
        "n_gpus": world, "fit_s": t_fit / 1e3, "broadcast_s": t_bc / 1e3, "broadcast_GB": state_bytes / 1e9,
        "predict_s": t_pred / 1e3, "gather_s": t_g / 1e3, "fit_predict_s": total / 1e3,
        "predict_TFLOPps_aggregate": (float(n) * n * 2 * MG) / (t_pred / 1e3) / 1e12,
        "mean_abs_max": float(torch.cat(mu).abs().max()), "var_min": float(torch.cat(vv).min()),
        "var_max": float(torch.cat(vv).max()), "gathered_points": int(mu[0].numel())}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
