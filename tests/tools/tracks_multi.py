"""Per-drifter track interpolation sharded over ranks (drifter n on rank n % world): every rank must
end with the same result as a single process.  Usage: torchrun --nproc-per-node 2 tests/tools/tracks_multi.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_track_interpolation import _drifters          # noqa: E402
from gp2d_b200 import laser_io_methods as lio            # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
data, _ = _drifters(n=12, seed=4, days=2.0)
torch.cuda.synchronize()
t0 = time.time()
tr = lio.interp_kriging(data, dt=900, period=2.0, optimize=True, max_iters=30, parallel=4)
torch.cuda.synchronize()
t1 = time.time()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()          # single-process reference below
    os.environ.pop("RANK", None)
# same fleet in this process alone (world() falls back to (0, 1) once the group is gone)
one = lio.interp_kriging(data, dt=900, period=2.0, optimize=True, max_iters=30, parallel=4)
same = all(np.array_equal(getattr(tr, k), getattr(one, k), equal_nan=True)
           for k in ("lon", "lat", "pos_varLon", "pos_varLat", "u", "v", "lenLon", "noiseLat", "n_samples"))
print("rank %d/%d: %d drifters, sharded run %.2f s, equals the single-process result: %s" % (rank, world, len(data), t1 - t0, same))
sys.exit(0 if same else 1)
