#!/usr/bin/env python
"""Headline benchmark: fit + predict seconds per snapshot (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--skip-configs2] [--skip-batched]

Headline workload at any N: BASELINE.json configs[1] -- one LASER-style snapshot, N_obs = 2000 drifter
observations (4000 x 4000 fp64 covariance), 320 x 320 = 102 400-point prediction grid,
curl-free + divergence-free kernel with l_df != l_cf (two exponentials per pair).  A step is
one fit + predict of one snapshot: covariance build, Cholesky / L^-1, alpha + LML, fused
K* / mean / variance.  With N > 1 GPUs every rank krigs its own independent snapshots
(weak scaling, no data-path collective; SURVEY.md §8e axis 3) and the value is
max-over-ranks time / total snapshots.

One JSON line on stdout (rank 0).  `value` times the device-resident path with CUDA events;
`e2e` times the call a user of the reference makes -- gp2d_fit_predict_host, pageable numpy arrays in
and out, host<->device copies inside the timed region.

The same line carries, measured once per invocation (not repeated --steps times):
  configs2  BASELINE.json's Target config: ONE factorisation of N = 16 384 observations on rank 0, its
            predict state broadcast over NCCL, the 1 000 000-point grid sharded over the --gpus ranks,
            mean / variance all-gathered (strong scaling: SURVEY.md §8e axis 1; krig.py:539-557);
  restarts  configs[3]: log-marginal-likelihood + gradient, 64 restarts over 8 GPUs (8 per GPU), the
            restarts of a rank advancing in lock step through gp2d_lml_grad_batched;
  snapshots configs[4]: 512 snapshots of N = 8192 over 8 GPUs (64 per GPU) through gp2d_fit_batched.

--impl reference: the oracle port of the reference's numpy path on the host cores (the reference
itself is Python 2 + GPy and cannot be imported here).  It never imports the product
package, so libgp2d.so is not mapped into that process.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBS = 2000
GRID = (320, 320)
THETA = (1.3, 3.1, 0.2)
NOISE = 0.05
CPU_GRID_FRACTION = 0.10          # share of the grid a bounded CPU step predicts (all of it for <= 3 steps)
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed
# ncu captures named in roofline.traffic_source (int8-sliced kernel) / roofline_fp64_kernel.traffic_source
NCU_PREDICT_I8_DRAM_BYTES = 47.294351e9 + 5.190785e9
NCU_PREDICT_I8_SOURCE = "profiles/r02m_predict_i8_kernel.md"
NCU_PREDICT_DRAM_BYTES = 108.822557e9 + 6.737815e9
NCU_PREDICT_SOURCE = "profiles/r01h_predict_kernel.md"
I8_SLICES = 6                     # slice count gp2d_fit picks at the conditioning of configs[1] / configs[2] / configs[4]
I8_PRODUCTS = I8_SLICES * (I8_SLICES + 1) // 2     # int8 x int8 MACs per fp64 MAC of the predictive product
# int8 issue rate of tcgen05.mma kind::i8 measured on this pool (tools/umma_probe2.cu, M=128 N=256 SS, all SMs)
I8_PROBE_TOPS = 4279.4
METRIC = "fit_predict_seconds_per_snapshot"
WORKLOAD = ("configs[1]: single LASER-style snapshot, N=2000 obs (4000x4000 fp64 covariance), "
            "320x320=102400-point grid, curl-free+div-free SE kernel theta=(1.3,3.1,0.2), noise 0.05")
C2_N = 16384
C2_GRID = 1000000


def config():
    return {"workload": WORKLOAD, "n_obs": N_OBS, "n": 2 * N_OBS, "grid_points": GRID[0] * GRID[1],
            "theta": list(THETA), "noise": NOISE,
            "l2": "working set (A + L^-1 = 268 MB) exceeds the 126 MB L2; no explicit flush",
            "sharding": "independent snapshots per rank, no data-path collective"}


def load_synthetic():
    """2d-gp_b200/synthetic.py by file path: the reference arm must not import the product package
    (its __init__ loads libgp2d.so)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gp2d_synthetic", os.path.join(ROOT, "2d-gp_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------
def cpu_step(X, y, Xs_sample, M_total):
    """One bounded CPU step: full fit at N=2000 + predict on a sample of the grid; the predict time is
    extrapolated linearly in M (exactly linear: SURVEY.md §8d).  Returns (extrapolated, measured, fit, predict)."""
    from oracle import gp_oracle as orc
    t0 = time.perf_counter()
    f = orc.fit(X, y, *THETA, NOISE)
    t1 = time.perf_counter()
    orc.predict(X, f, *THETA, Xs_sample, chunk=2048)
    t2 = time.perf_counter()
    fit_s, pred_s = t1 - t0, t2 - t1
    return fit_s + pred_s * (M_total / Xs_sample.shape[0]), fit_s + pred_s, fit_s, pred_s


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_host():
    import platform
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    model = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return "%s, %d logical CPUs (%s)" % (model or platform.processor() or "cpu", os.cpu_count() or 1, platform.node())


def reexec_with_all_threads():
    """torchrun exports OMP_NUM_THREADS=1; BLAS reads it when it is loaded.  Start again with every
    host core before numpy is imported."""
    n = str(os.cpu_count() or 1)
    if os.environ.get("GP2D_BENCH_REEXEC") == "1":
        return
    if all(os.environ.get(k) == n for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")):
        return
    env = dict(os.environ)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        env[k] = n
    env["GP2D_BENCH_REEXEC"] = "1"
    sys.stdout.flush()
    os.execve(sys.executable, [sys.executable] + sys.argv, env)


def repo_shared_objects():
    """Shared objects of this repository mapped into the process (the reference arm must list none)."""
    found = set()
    try:
        with open("/proc/self/maps") as f:
            for ln in f:
                path = ln.rsplit(" ", 1)[-1].strip()
                if path.startswith(ROOT) and ".so" in os.path.basename(path):
                    found.add(os.path.relpath(path, ROOT))
    except OSError:
        pass
    return sorted(found)


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    reexec_with_all_threads()
    import numpy as np
    syn = load_synthetic()
    X, y = syn.drifter_snapshot(N_OBS, config_id=2, seed_offset=0)
    Xs = syn.prediction_grid(X, GRID[0], GRID[1])
    M = Xs.shape[0]
    frac = args.cpu_fraction if args.cpu_fraction else (1.0 if args.steps + args.warmup <= 3 else CPU_GRID_FRACTION)
    nsample = max(64, min(M, int(round(M * frac))))
    sample = Xs if nsample == M else Xs[np.random.default_rng(0).choice(M, nsample, replace=False)]
    for _ in range(args.warmup):
        cpu_step(X, y, sample, M)
    rows = [cpu_step(X, y, sample, M) for _ in range(args.steps)]
    val = float(np.mean([r[0] for r in rows]))
    measured = float(np.mean([r[1] for r in rows]))
    cores = cpu_threads()
    sample_txt = ("oracle port (numpy/scipy, OpenBLAS, %d threads) on %s: every step = full fit at N=2000 (%.2f s) + "
                  "predict on %d of %d grid points (%.2f s); value = fit + predict x %.1f (predict is exactly "
                  "linear in M), measured_s = what a step took" %
                  (cores, cpu_host(), float(np.mean([r[2] for r in rows])), nsample, M,
                   float(np.mean([r[3] for r in rows])), M / nsample))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": measured * 1e3, "measured_s": measured,
        "extrapolation_factor_on_predict": M / nsample, "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(),
        "cpu_baseline": {"value": val, "unit": "s", "cores": cores, "kind": "port", "sample": sample_txt,
                         "host": cpu_host()},
        "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "repo_shared_objects_mapped": repo_shared_objects(),
    }))


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        import numpy as np
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def potri_launches(nb, need_inv=True):
    if nb == 1:
        return 1
    n1 = nb // 2
    return potri_launches(n1, True) + 2 + potri_launches(nb - n1, need_inv) + (2 if need_inv else 0)


def i8_counters():
    """Slice products / stages / dense k-steps / launches of the int8 predictive kernel since the last call (device counters)."""
    import ctypes as C
    from gp2d_b200._lib import lib
    buf = (C.c_ulonglong * 4)()
    lib.gp2d_dbg_i8_counters.restype = C.c_int
    lib.gp2d_dbg_i8_counters.argtypes = [C.POINTER(C.c_ulonglong)]
    if lib.gp2d_dbg_i8_counters(buf) != 0:
        raise RuntimeError("gp2d_dbg_i8_counters failed")
    return [int(v) for v in buf]


def bf16_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        return 1590.0       # fallback stated in B200_PROFILING.md


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0       # fallback stated in B200_PROFILING.md


def run_configs2(torch, dist, gp, gdist, syn, rank, world, dev, peak_tf):
    """BASELINE.json's Target: fit + predict of N = 16 384 observations onto a 1 000 000-point grid.
    Rank 0 factorises once; the predict state (L^-1 tiles, alpha, X) is broadcast; the grid is sharded
    block-cyclically in whole column tiles (krig.py:539-557 predicts slice by slice; dist.shard_cyclic); mean /
    variance are all-gathered (in shard order here: the block only takes statistics of them).  CUDA-event times, max over ranks."""
    import numpy as np
    N, MG = C2_N, C2_GRID
    X, y = syn.drifter_snapshot(N, config_id=3)
    side = int(np.ceil(np.sqrt(MG)))
    Xs = syn.prediction_grid(X, side, side)[:MG]
    idx = gdist.shard_cyclic(MG, rank, world)       # block-cyclic share of the grid: tile costs vary along the grid
    mloc = int(idx.shape[0])

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def tmax(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    m = gp.HelmholtzGP(X, y, *THETA, NOISE)
    Xsd = gp.as_dev(Xs[idx])
    alpha = torch.empty(2 * N, dtype=torch.float64, device=dev)
    # warm-up of the predict path (module load, K* panel scratch) on a zeroed state: no rank holds a
    # factorisation before the timed region, so the ranks > 0 can only get theirs from the broadcast
    m.ws.zero_()
    m.fitted = True
    m.predict(Xsd[:64])
    m.fitted = False
    torch.cuda.synchronize()
    if world > 1:                # NCCL connection set-up outside the timed region
        dist.broadcast(torch.zeros(1 << 20, dtype=torch.uint8, device=dev), src=0)
        dist.barrier()
    torch.cuda.synchronize()
    e0 = ev()
    if rank == 0:
        m.fit_async(alpha_out=alpha)
    e1 = ev()
    gdist.broadcast_fit(m, src=0)
    i8_counters()                 # clear (synchronises: the broadcast has landed on this rank)
    e2 = ev()
    mean, var = m.predict(Xsd)
    e3 = ev()
    mu = [gdist.gather_concat(mean[:mloc].contiguous()), gdist.gather_concat(mean[mloc:].contiguous())]
    vv = [gdist.gather_concat(var[:mloc].contiguous()), gdist.gather_concat(var[mloc:].contiguous())]
    e4 = ev()
    torch.cuda.synchronize()
    # the receiving ranks sit in the broadcast while rank 0 is still factorising: the transfer itself is
    # rank 0's own broadcast time
    cnt = i8_counters()            # this rank's grid slice: slice products issued, stages, dense k-steps
    csum = torch.tensor([float(cnt[0]), float(cnt[2])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(csum)
    t_fit, t_pred, t_g = tmax(e0.elapsed_time(e1)), tmax(e2.elapsed_time(e3)), tmax(e3.elapsed_time(e4))
    t_bc = tmax(e1.elapsed_time(e2) if rank == 0 else 0.0)
    total = tmax(e0.elapsed_time(e4))
    n = 2 * N
    out = None
    if rank == 0:
        # in-run parity: GP identity at observation points, mean(X_i) = y_i - (noise) alpha_i
        idx = np.linspace(0, N - 1, 512).astype(np.int64)
        mo, vo = m.predict(X[idx])
        a = alpha.cpu().numpy()
        ident = np.concatenate([y[idx] - NOISE * a[idx], y[N + idx] - NOISE * a[N + idx]])
        resid = float(np.abs(mo.cpu().numpy() - ident).max() / np.abs(y).max())
        state_bytes = m.predict_state().numel()
        flops_pred = float(n) * n * 2 * MG + 2.0 * n * 2 * MG
        flops_fit = 2.0 * float(n) ** 3 / 3.0
        kss = THETA[2] / THETA[0] ** 2 + (1 - THETA[2]) / THETA[1] ** 2
        allv = torch.cat(vv)
        out = {
            "workload": "configs[2]: N=%d obs (n=%d fp64 Cholesky), %d-point grid sharded over %d GPU(s); ONE "
                        "factorisation on rank 0, predict state broadcast, grid dealt block-cyclically (320-point blocks) to the ranks, all_gather" % (N, n, MG, world),
            "scaling": "strong", "n_gpus": world, "fit_s": t_fit / 1e3, "broadcast_s": t_bc / 1e3,
            "broadcast_GB": state_bytes / 1e9,
            "broadcast_GBps": (state_bytes / 1e9) / (t_bc / 1e3) if world > 1 and t_bc > 0 else None,
            "predict_s": t_pred / 1e3, "gather_s": t_g / 1e3, "total_s": total / 1e3,
            "fit_TFLOPps": flops_fit / (t_fit / 1e3) / 1e12, "fit_frac_of_fp64_peak": flops_fit / (t_fit / 1e3) / 1e12 / peak_tf,
            "predict_kernel": "int8-sliced tcgen05 kernel, %d slices (fp64 DMMA kernel when the fit state says 0): slices=%d"
                              % (I8_SLICES, int(m.ws[-256:].view(torch.int32)[2].item())),
            "predict_TFLOPps_aggregate": flops_pred / (t_pred / 1e3) / 1e12,
            "predict_TFLOPps_note": "fp64-equivalent: the algorithmic flops n^2 m + 2 n m of the fp64 product over the time",
            "predict_x_fp64_pipe_peak": flops_pred / (t_pred / 1e3) / 1e12 / (world * peak_tf),
            "predict_int8_TOPps_aggregate": 2.0 * 128 * 80 * 32 * float(csum[0].item()) / (t_pred / 1e3) / 1e12,
            "predict_frac_of_int8_peak": 2.0 * 128 * 80 * 32 * float(csum[0].item()) / (t_pred / 1e3) / 1e12 / (world * 2.0 * bf16_peak()),
            "predict_int8_note": "int8 ops of the slice products the kernel issued (device counters); identically zero digit slices "
                                 "are skipped: %.3f of the dense schedule's %d products per k-step were issued"
                                 % (float(csum[0].item()) / max(1.0, I8_PRODUCTS * float(csum[1].item())), I8_PRODUCTS),
            "total_TFLOPps_aggregate": (flops_pred + flops_fit) / (total / 1e3) / 1e12,
            "total_x_fp64_pipe_peak": (flops_pred + flops_fit) / (total / 1e3) / 1e12 / (world * peak_tf),
            "unsharded_share_of_total": (t_fit + t_bc) / total,
            "parity": {"gp_identity_residual_rel": resid, "var_min": float(allv.min()), "var_max": float(allv.max()),
                       "var_upper_bound_kss": kss, "gathered_points": int(mu[0].numel()), "info": int(m._info.item()),
                       "check": "mean at 512 observation points vs y - noise*alpha (rel. to max|y|); 0 <= var <= k**"},
        }
    del m, Xsd, mean, var, mu, vv
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import gp2d_b200 as gp
    from gp2d_b200 import dist as gdist
    from gp2d_b200 import synthetic as syn
    from gp2d_b200._lib import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the GPU path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def snapshot_and_grid(seed_offset):
        X, y = syn.drifter_snapshot(N_OBS, config_id=2, seed_offset=seed_offset)
        return X, y, syn.prediction_grid(X, GRID[0], GRID[1])

    K, W = args.steps, args.warmup
    # each rank owns its own stream of snapshots (seed offset = global snapshot id)
    snaps = [snapshot_and_grid(rank * (K + W) + i) for i in range(min(K + W, 4))]
    M = snaps[0][2].shape[0]
    n = 2 * N_OBS
    npad = (n + 127) // 128 * 128
    dsn = [(gp.as_dev(X), gp.as_dev(y), gp.as_dev(Xs)) for (X, y, Xs) in snaps]
    model = gp.HelmholtzGP(dsn[0][0], dsn[0][1], *THETA, NOISE)
    mean = torch.empty(2 * M, dtype=torch.float64, device=dev)
    var = torch.empty(2 * M, dtype=torch.float64, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    pred_ev = [(ev(), ev()) for _ in range(K)]

    def step(i, timed_idx=None):
        Xd, yd, Xsd = dsn[i % len(dsn)]
        model.X, model.y = Xd, yd
        model.fit_async()
        if timed_idx is not None:
            pred_ev[timed_idx][0].record()
        model.predict(Xsd, out_mean=mean, out_var=var)
        if timed_idx is not None:
            pred_ev[timed_idx][1].record()

    for i in range(W):
        step(i)
    i8_counters()                             # clear: count the slice products of the K timed launches only
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(K):
        step(W + i, i)
    e1.record()
    barrier()
    clocks = sampler.stop()
    t_dev = e0.elapsed_time(e1) * 1e-3
    info = int(model._info.item())
    tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt.item())
    value = t_max / (K * world)
    pred_ms = float(np.mean([a.elapsed_time(b) for a, b in pred_ev]))
    i8_cnt = i8_counters()                    # slice products, stages, k-steps of the dense schedule, launches

    # ---- the fp64 tensor-pipe kernel on the same snapshot (GP2D_OPT_PREDICT_I8 = 1): the kernel the int8 path
    # replaces at this conditioning and the only path beyond it; kept in the line so both rooflines are visible ----
    i8_slices = int(model.ws[-256:].view(torch.int32)[2].item())
    gp.set_predict_i8(1)
    model.fit_async()
    model.predict(dsn[0][2], out_mean=mean, out_var=var)
    f0, f1 = ev(), ev()
    f0.record()
    for _ in range(3):
        model.predict(dsn[0][2], out_mean=mean, out_var=var)
    f1.record()
    torch.cuda.synchronize()
    pred_ms_fp64 = f0.elapsed_time(f1) / 3
    gp.set_predict_i8(0)
    model.fit_async()
    torch.cuda.synchronize()

    # ---- e2e: the call a user of the reference makes.  gp2d_fit_predict_host takes pageable numpy
    # arrays (what GP_laser.simLaser / GPRegression.predict hand over), copies them to the device,
    # fits, predicts and copies mean / variance / LML back; all of it inside the timed region. ------
    def e2e_step(i):
        Xh, yh, Xsh = snaps[i % len(snaps)]
        return gp.fit_predict_host(Xh, yh, *THETA, NOISE, Xsh)

    for i in range(max(1, min(W, 3))):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        mu_h, var_h, lml_h = e2e_step(W + i)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0          # host clock: the call synchronises before it returns
    barrier()
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = float(te.item()) / (K * world)
    h2d = 8 * (2 * N_OBS + 2 * N_OBS + 2 * M)
    d2h = 8 * (2 * M + 2 * M) + 8 + 4
    lib.gp2d_host_release()

    # the GPy-style route on the same snapshot: GPRegression(X, Y, kern) evaluates the likelihood AND its
    # gradient in the constructor (as GPy does), then predict(Xnew) with numpy in / out
    e2e_gpy = None
    if rank == 0:
        from gp2d_b200 import models, myKernel
        Xh, yh, Xsh = snaps[0]
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mm = models.GPRegression(Xh, yh[:, None], myKernel.myKernel(2, [0, 1], *THETA), noise_var=NOISE, jitter=0.0)
            mm.predict(Xsh)
            e2e_gpy = time.perf_counter() - t0
        del mm

    # ---- live FP64 tensor-pipe ceiling (every rank: configs2 needs it on rank 0 only) ---------------
    import ctypes as C
    st = torch.cuda.current_stream().cuda_stream
    scratch = torch.zeros(8, dtype=torch.float64, device=dev)
    ctas, iters = 148 * 2, 20000
    lib.gp2d_dbg_fp64_peak.restype = C.c_int
    lib.gp2d_dbg_fp64_peak.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    for _ in range(2):
        lib.gp2d_dbg_fp64_peak(iters, ctas, scratch.data_ptr(), st)
    p0, p1 = ev(), ev()
    p0.record()
    lib.gp2d_dbg_fp64_peak(iters, ctas, scratch.data_ptr(), st)
    p1.record()
    torch.cuda.synchronize()
    peak_tf = ctas * 8 * iters * 16 * 512 / (p0.elapsed_time(p1) * 1e-3) / 1e12

    out = None
    stages = targets = cpu = None
    if rank == 0:
        def timeit(fn, reps=5):
            fn(); torch.cuda.synchronize()
            a, b = ev(), ev()
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) * 1e-3 / reps

        Kfull = torch.empty((n, n), dtype=torch.float64, device=dev)
        t_build = timeit(lambda: gp.kernel_K(dsn[0][0], None, *THETA, diag_add=NOISE, out=Kfull))
        nb_ws = lib.gp2d_potrf_workspace_bytes(n)
        ws = torch.empty(nb_ws, dtype=torch.uint8, device=dev)
        info_t = torch.zeros(1, dtype=torch.int32, device=dev)
        Kwork = torch.empty_like(Kfull)

        def do_potrf():
            Kwork.copy_(Kfull)
            lib.gp2d_potrf(Kwork.data_ptr(), n, n, ws.data_ptr(), nb_ws, info_t.data_ptr(), st)
        t_copy = timeit(lambda: Kwork.copy_(Kfull))
        t_potrf = timeit(do_potrf) - t_copy
        t_fit = timeit(lambda: model.fit_async())
        stages = {
            "fit_ms": t_fit * 1e3, "predict_ms": pred_ms,
            "kernel_build_GBps": 8.0 * n * n / t_build / 1e9,
            "kernel_build_frac_of_hbm_peak": (8.0 * n * n / t_build / 1e9) / hbm_peak(),
            "cholesky_TFLOPps": (n ** 3 / 3.0) / t_potrf / 1e12,
            "cholesky_frac_of_fp64_peak": (n ** 3 / 3.0) / t_potrf / 1e12 / peak_tf,
            "potrf_ms": t_potrf * 1e3, "build_ms": t_build * 1e3,
            "note": "at configs[1] size both probes are latency-bound (a 128 MB build is one 25 us launch; the "
                    "Cholesky of a 4000 x 4000 matrix is a chain of 32 diagonal blocks and ~120 small GEMMs); the "
                    "stage targets of BASELINE.json are quoted at N=16384, see targets_at_N16384",
        }

        # ---- the two stage targets BASELINE.json quotes at configs[2] size (N=16384, n=32768) --------
        del Kfull, Kwork, ws
        torch.cuda.empty_cache()
        Nb = C2_N
        nbig = 2 * Nb
        Xb = gp.as_dev(syn.drifter_snapshot(Nb, config_id=3)[0])
        Kb = torch.empty((nbig, nbig), dtype=torch.float64, device=dev)
        t_build_big = timeit(lambda: gp.kernel_K(Xb, None, *THETA, diag_add=NOISE, out=Kb), reps=3)
        nb_ws = lib.gp2d_potrf_workspace_bytes(nbig)
        ws = torch.empty(nb_ws, dtype=torch.uint8, device=dev)

        def do_potrf_big():
            gp.kernel_K(Xb, None, *THETA, diag_add=NOISE, out=Kb)
            lib.gp2d_potrf(Kb.data_ptr(), nbig, nbig, ws.data_ptr(), nb_ws, info_t.data_ptr(), st)
        t_potrf_big = timeit(do_potrf_big, reps=2) - t_build_big
        info_big = int(info_t.item())
        del Kb, ws, Xb
        torch.cuda.empty_cache()
        targets = {
            "what": "BASELINE.json north-star stage targets, measured at configs[2] size (n=32768) on this GPU",
            "kernel_build_GBps": 8.0 * nbig * nbig / t_build_big / 1e9,
            "kernel_build_frac_of_hbm_peak": (8.0 * nbig * nbig / t_build_big / 1e9) / hbm_peak(),
            "kernel_build_target_frac": 0.70,
            "cholesky_TFLOPps": (nbig ** 3 / 3.0) / t_potrf_big / 1e12,
            "cholesky_frac_of_fp64_peak": (nbig ** 3 / 3.0) / t_potrf_big / 1e12 / peak_tf,
            "cholesky_target_frac": 0.60, "potrf_ms": t_potrf_big * 1e3, "build_ms": t_build_big * 1e3,
            "potrf_info": info_big,
        }

        # ---- CPU baseline (rank 0, N=1 only: the contract's bounded sample) ---------------------
        if world == 1:
            from threadpoolctl import threadpool_limits
            X, y, Xs = snaps[0]
            nsample = int(round(M * CPU_GRID_FRACTION))
            sample = Xs[np.random.default_rng(0).choice(M, nsample, replace=False)]
            with threadpool_limits(limits=os.cpu_count() or 1):
                cpu_val, cpu_meas, cpu_fit, cpu_pred = cpu_step(X, y, sample, M)
                cores = cpu_threads()
            cpu = {"value": cpu_val, "unit": "s", "cores": cores, "kind": "port", "host": cpu_host(), "measured_s": cpu_meas,
                   "sample": "oracle (numpy/scipy): full fit at N=2000 (%.2f s) + predict on %d of %d grid points "
                             "(%.2f s) extrapolated linearly in M" % (cpu_fit, nsample, M, cpu_pred)}

    del model, dsn, mean, var
    torch.cuda.empty_cache()

    # ---- BASELINE.json's Target config and its multi-GPU split, once per invocation ------------------
    c2 = None
    if not args.skip_configs2:
        barrier()
        c2 = run_configs2(torch, dist, gp, gdist, syn, rank, world, dev, peak_tf)

    batched = None
    if not args.skip_batched:
        barrier()
        try:
            import bench_batched
        except ImportError:
            bench_batched = None
        if bench_batched is not None:
            batched = bench_batched.run(torch, dist, gp, gdist, syn, rank, world, dev, peak_tf)

    if rank == 0:
        m_cols = 2 * M
        flops_pred = float(n) * n * m_cols + 2.0 * n * m_cols
        ach = flops_pred / (pred_ms * 1e-3) / 1e12
        sm_max_mhz = clocks.get("sm_max_mhz") or 1965.0
        derived = 148 * 64 * 2 * sm_max_mhz * 1e6 / 1e12
        ach64 = flops_pred / (pred_ms_fp64 * 1e-3) / 1e12
        peak_src = ("FP64 DMMA.8x8x4 register-resident loop measured live in this run (MEASURED_PEAKS.json has no fp64 figure); "
                    "derived ceiling 148 SMs x 64 FP64 FMA/clk/SM x 2 flop x %.0f MHz = %.2f TFLOP/s" % (sm_max_mhz, derived))
        roofline_fp64 = {
            "kernel": "predict_kernel (fused K* generation + Z K*^T on DMMA + mean/variance), GP2D_OPT_PREDICT_I8 = 1",
            "bound": "tensor", "achieved": ach64, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach64 / peak_tf,
            "traffic": NCU_PREDICT_DRAM_BYTES, "traffic_unit": "bytes", "traffic_source": NCU_PREDICT_SOURCE,
            "peak_source": peak_src, "peak_derived": derived, "algorithmic_flops_per_launch": flops_pred, "ms_per_launch": pred_ms_fp64,
        }
        if i8_slices:
            nprod = i8_slices * (i8_slices + 1) // 2
            i8_peak = 2.0 * bf16_peak()
            nc = 80 if i8_slices == 6 else 64                     # accumulator columns of the slice-count variant
            launches = max(1, int(i8_cnt[3]))
            ops_exec = 2.0 * 128 * nc * 32 * float(i8_cnt[0]) / launches     # int8 ops of the MMAs actually issued, per launch
            ops_dense = 2.0 * 128 * nc * 32 * nprod * float(i8_cnt[2]) / launches
            ach8 = ops_exec / (pred_ms * 1e-3) / 1e12
            roofline = {
                "kernel": "predict_i8_kernel<%d> (fused K* digit-slice generation + %d tcgen05.mma kind::i8 slice products per "
                          "k-step into TMEM + fp64 recombination, mean/variance)" % (i8_slices, nprod),
                "bound": "tensor", "achieved": ach8, "peak": i8_peak, "unit": "TOP/s", "frac": ach8 / i8_peak,
                "traffic": NCU_PREDICT_I8_DRAM_BYTES, "traffic_unit": "bytes", "traffic_source": NCU_PREDICT_I8_SOURCE,
                "peak_source": "int8 dense tensor rate = 2 x the measured bf16 figure of MEASURED_PEAKS.json (%.1f TFLOP/s; the file has "
                               "no int8 entry; nominal 4500 TOP/s); this pool's own tcgen05 kind::i8 issue-rate probe "
                               "(tools/umma_probe2.cu, N = 256) reads %.0f TOP/s" % (bf16_peak(), I8_PROBE_TOPS),
                "frac_of_probe_peak": ach8 / I8_PROBE_TOPS,
                "algorithmic_ops_per_launch": ops_exec,
                "algorithmic_ops_note": "int8 ops of the tcgen05.mma instructions the kernel issued (counted on the device, "
                                        "gp2d_dbg_i8_counters): of the %d slice pairs (i, j), i + j < %d, per k-step of the padded "
                                        "lower-triangular product, those whose digit slices are not identically zero -- %.3f of the "
                                        "dense schedule here (%.4g ops; the unpadded n^2 m + 2 n m fp64 flops of SURVEY.md 8d x %d "
                                        "x 1 int8 op per flop would be %.4g)" % (nprod, i8_slices, ops_exec / max(ops_dense, 1.0),
                                                                                 ops_dense, nprod, nprod * flops_pred),
                "slice_products_issued_per_launch": float(i8_cnt[0]) / launches,
                "stages_issued_per_launch": float(i8_cnt[1]) / launches,
                "ksteps_dense_per_launch": float(i8_cnt[2]) / launches,
                "dense_schedule_equivalent": {"achieved": ops_dense / (pred_ms * 1e-3) / 1e12, "unit": "TOP/s",
                                              "note": "the same time against every slice product of the dense schedule (zero "
                                                      "slices included): what a kernel without the skip would have to sustain"},
                "ms_per_launch": pred_ms,
                "fp64_equivalent": {"achieved": ach, "unit": "TFLOP/s", "fp64_pipe_peak": peak_tf, "x_fp64_pipe_peak": ach / peak_tf,
                                    "speedup_over_fp64_kernel": pred_ms_fp64 / pred_ms},
                "structural_bound": "measured with tools/umma_probe4.cu on this pool: a kind::i8 MMA at M = 128 costs at least 47 clocks "
                                    "whatever N (and whether A comes from shared memory or TMEM), 54-55 at N = 80 in SS mode against "
                                    "40 clocks of tensor time (TMEM holds 6 accumulators x 80 columns, so N cannot grow), and the pipe "
                                    "queues nothing behind the instruction it executes: 0.73 of the int8 peak is the ceiling of this "
                                    "tile shape; copies into shared memory, the K* generators and the epilogue share the SM with it",
            }
        else:
            roofline = roofline_fp64
        out = {
            "metric": METRIC, "value": value, "unit": "s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_max / K * 1e3, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(),
            "dtype_note": "fp64 throughout; the predictive product Z K* runs as exact int8 digit-slice products (int32 accumulation in "
                          "TMEM) recombined in fp64 when the fit's conditioning allows it (slices=%d here), on the FP64 tensor pipe otherwise" % i8_slices,
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": "gp2d_fit_predict_host (C ABI, pageable numpy arrays in and out, device buffer cached per thread)",
                    "gpregression_construct_plus_predict_s": e2e_gpy,
                    "gpregression_note": "models.GPRegression(X, Y, myKernel).predict(Xnew): the constructor evaluates LML "
                                         "and its gradient like GPy's, then predict refits and predicts"},
            # spatial order (sort, gather), build, potri tree, pack, digit slices (row scale, quantise, gate), alpha/LML,
            # predict (the int8 kernel of each slice count + the fp64 kernel: the fit state selects one on the device,
            # the others return at once)
            "gpu_launches": K * (2 + 1 + potri_launches(npad // 128) + 1 + 3 + 5 + 3),
            "roofline": roofline, "roofline_fp64_kernel": roofline_fp64,
            "stages": stages, "targets_at_N16384": targets, "info": info,
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        if c2 is not None:
            out["configs2"] = c2
        if batched is not None:
            out.update(batched)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-configs2", action="store_true", help="leave out the N=16384 / 1M-grid Target block (~65 s on one GPU)")
    ap.add_argument("--skip-batched", action="store_true", help="leave out the restarts / snapshots blocks")
    ap.add_argument("--cpu-fraction", type=float, default=0.0,
                    help="reference arm: share of the grid predicted per step (default 0.10, 1.0 for <= 3 steps)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
