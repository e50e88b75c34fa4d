#!/usr/bin/env python
"""Headline benchmark: fit + predict seconds per snapshot (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at any N: BASELINE.json configs[1] -- one LASER-style snapshot, N_obs = 2000 drifter
observations (4000 x 4000 fp64 covariance), 320 x 320 = 102 400-point prediction grid,
curl-free + divergence-free kernel with l_df != l_cf (two exponentials per pair).  A step is
one fit + predict of one snapshot: covariance build, Cholesky / L^-1, alpha + LML, fused
K* / mean / variance.  With N > 1 GPUs every rank krigs its own independent snapshots
(weak scaling, no data-path collective; SURVEY.md §8e axis 3) and the value is
max-over-ranks time / total snapshots.

One JSON line on stdout (rank 0).  `value` times the device-resident path with CUDA events;
`e2e` times the public host-array API (pinned host buffers, H2D + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBS = 2000
GRID = (320, 320)
THETA = (1.3, 3.1, 0.2)
NOISE = 0.05
CPU_GRID_SAMPLE = 1024
NCU_PREDICT_DRAM_BYTES = 108.822557e9 + 6.737815e9
METRIC = "fit_predict_seconds_per_snapshot"
WORKLOAD = ("configs[1]: single LASER-style snapshot, N=2000 obs (4000x4000 fp64 covariance), "
            "320x320=102400-point grid, curl-free+div-free SE kernel theta=(1.3,3.1,0.2), noise 0.05")


def config():
    return {"workload": WORKLOAD, "n_obs": N_OBS, "n": 2 * N_OBS, "grid_points": GRID[0] * GRID[1],
            "theta": list(THETA), "noise": NOISE,
            "l2": "working set (A + L^-1 = 268 MB) exceeds the 126 MB L2; no explicit flush",
            "sharding": "independent snapshots per rank, no data-path collective"}


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------
def cpu_step(X, y, Xs_sample, M_total):
    """One bounded CPU sample: full fit at N=2000, predict on CPU_GRID_SAMPLE grid points,
    predict time extrapolated linearly in M (exactly linear: SURVEY.md §8d)."""
    from oracle import gp_oracle as orc
    t0 = time.perf_counter()
    f = orc.fit(X, y, *THETA, NOISE)
    t1 = time.perf_counter()
    orc.predict(X, f, *THETA, Xs_sample, chunk=CPU_GRID_SAMPLE)
    t2 = time.perf_counter()
    return (t1 - t0) + (t2 - t1) * (M_total / Xs_sample.shape[0]), (t1 - t0), (t2 - t1)


def snapshot_and_grid(seed_offset):
    from gp2d_b200 import synthetic
    X, y = synthetic.drifter_snapshot(N_OBS, config_id=2, seed_offset=seed_offset)
    return X, y, synthetic.prediction_grid(X, GRID[0], GRID[1])


def snapshot_big(N):
    from gp2d_b200 import synthetic
    return synthetic.drifter_snapshot(N, config_id=3)[0]


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, y, Xs = snapshot_and_grid(0)
    M = Xs.shape[0]
    sample = Xs[np.random.default_rng(0).choice(M, CPU_GRID_SAMPLE, replace=False)]
    for _ in range(args.warmup):
        cpu_step(X, y, sample, M)
    times = [cpu_step(X, y, sample, M)[0] for _ in range(args.steps)]
    val = float(np.mean(times))
    cores = cpu_threads()
    sample_txt = ("oracle port (numpy/scipy, OpenBLAS): full fit at N=2000 + predict on %d of %d grid "
                  "points, predict time extrapolated linearly in M" % (CPU_GRID_SAMPLE, M))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": val * 1e3, "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(),
        "cpu_baseline": {"value": val, "unit": "s", "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gp2d_b200 as gp
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

    # ---- e2e: public host-array API, pinned host buffers, H2D + D2H every step ------------
    hsn = [tuple(torch.from_numpy(a).pin_memory() for a in s) for s in snaps]
    h_mean = torch.empty(2 * M, dtype=torch.float64).pin_memory()
    h_var = torch.empty(2 * M, dtype=torch.float64).pin_memory()

    def e2e_step(i):
        Xh, yh, Xsh = hsn[i % len(hsn)]
        m = model
        m.X = Xh.to(dev, non_blocking=True)
        m.y = yh.to(dev, non_blocking=True)
        lml = m.fit()                                  # D2H of LML + info (synchronises)
        mu, vv = m.predict(Xsh.to(dev, non_blocking=True), out_mean=mean, out_var=var)
        h_mean.copy_(mu, non_blocking=True)
        h_var.copy_(vv, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return lml

    for i in range(max(1, min(W, 3))):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    ee0, ee1 = ev(), ev()
    ee0.record()
    for i in range(K):
        e2e_step(W + i)
    ee1.record()
    barrier()
    t_e2e_host = time.perf_counter() - t0
    t_e2e = max(ee0.elapsed_time(ee1) * 1e-3, t_e2e_host)
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = float(te.item()) / (K * world)
    h2d = 8 * (2 * N_OBS + 2 * N_OBS + 2 * M)
    d2h = 8 * (2 * M + 2 * M) + 8 + 4

    out = None
    if rank == 0:
        # ---- live FP64 tensor-pipe ceiling + stage numbers (outside the timed regions) ------
        scratch = torch.zeros(8, dtype=torch.float64, device=dev)
        ctas, iters = 148 * 2, 20000
        st = torch.cuda.current_stream().cuda_stream
        import ctypes as C
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

        # ---- the two stage targets BASELINE.json quotes at configs[2] size (N=16384, n=32768) --------
        del Kfull, Kwork, ws
        torch.cuda.empty_cache()
        Nb = 16384
        nbig = 2 * Nb
        Xb = gp.as_dev(snapshot_big(Nb))
        Kb = torch.empty((nbig, nbig), dtype=torch.float64, device=dev)
        t_build_big = timeit(lambda: gp.kernel_K(Xb, None, *THETA, diag_add=NOISE, out=Kb), reps=3)
        nb_ws = lib.gp2d_potrf_workspace_bytes(nbig)
        ws = torch.empty(nb_ws, dtype=torch.uint8, device=dev)

        def do_potrf_big():
            gp.kernel_K(Xb, None, *THETA, diag_add=NOISE, out=Kb)
            lib.gp2d_potrf(Kb.data_ptr(), nbig, nbig, ws.data_ptr(), nb_ws, info_t.data_ptr(), st)
        t_potrf_big = timeit(do_potrf_big, reps=2) - t_build_big
        info_big = int(info_t.item())
        del Kb, ws
        torch.cuda.empty_cache()

        m_cols = 2 * M
        flops_pred = float(n) * n * m_cols + 2.0 * n * m_cols
        ach = flops_pred / (pred_ms * 1e-3) / 1e12

        # ---- CPU baseline (rank 0, bounded sample) -----------------------------------------
        X, y, Xs = snaps[0]
        sample = Xs[np.random.default_rng(0).choice(M, CPU_GRID_SAMPLE, replace=False)]
        cpu_step(X, y, sample, M)
        cpu_val, cpu_fit, cpu_pred = cpu_step(X, y, sample, M)

        out = {
            "metric": METRIC, "value": value, "unit": "s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_max / K * 1e3, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": K * (1 + potri_launches(npad // 128) + 1 + 5 + 1),   # build, potri tree, pack, alpha/LML, predict
            "roofline": {
                "kernel": "predict_kernel (fused K* generation + Z K*^T DMMA + mean/variance)",
                "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture
                # profiles/r01h_predict_kernel.md (re-reads of the per-CTA K* panel; 14.6 % of DRAM peak)
                "traffic": NCU_PREDICT_DRAM_BYTES, "traffic_unit": "bytes",
                "peak_source": "FP64 DMMA.8x8x4 register-resident loop measured live in this run "
                               "(MEASURED_PEAKS.json has no fp64 figure; HGX B200 datasheet: 37 TFLOP/s)",
                "algorithmic_flops_per_launch": flops_pred, "ms_per_launch": pred_ms,
            },
            "cpu_baseline": {
                "value": cpu_val, "unit": "s", "cores": cpu_threads(), "kind": "port",
                "sample": "oracle (numpy/scipy): full fit at N=2000 (%.2f s) + predict on %d of %d grid points "
                          "(%.2f s) extrapolated linearly in M" % (cpu_fit, CPU_GRID_SAMPLE, M, cpu_pred)},
            "stages": {
                "fit_ms": t_fit * 1e3, "predict_ms": pred_ms,
                "kernel_build_GBps": 8.0 * n * n / t_build / 1e9,
                "kernel_build_frac_of_hbm_peak": (8.0 * n * n / t_build / 1e9) / hbm_peak(),
                "cholesky_TFLOPps": (n ** 3 / 3.0) / t_potrf / 1e12,
                "cholesky_frac_of_fp64_peak": (n ** 3 / 3.0) / t_potrf / 1e12 / peak_tf,
                "potrf_ms": t_potrf * 1e3, "build_ms": t_build * 1e3,
                "note": "at configs[1] size both probes are latency-bound (a 128 MB build is one 25 us launch; the "
                        "Cholesky of a 4000 x 4000 matrix is a chain of 32 diagonal blocks and ~120 small GEMMs); the "
                        "stage targets of BASELINE.json are quoted at N=16384, see targets_at_N16384",
            },
            "targets_at_N16384": {
                "what": "BASELINE.json north-star stage targets, measured at configs[2] size (n=32768) on this GPU",
                "kernel_build_GBps": 8.0 * nbig * nbig / t_build_big / 1e9,
                "kernel_build_frac_of_hbm_peak": (8.0 * nbig * nbig / t_build_big / 1e9) / hbm_peak(),
                "kernel_build_target_frac": 0.70,
                "cholesky_TFLOPps": (nbig ** 3 / 3.0) / t_potrf_big / 1e12,
                "cholesky_frac_of_fp64_peak": (nbig ** 3 / 3.0) / t_potrf_big / 1e12 / peak_tf,
                "cholesky_target_frac": 0.60, "potrf_ms": t_potrf_big * 1e3, "build_ms": t_build_big * 1e3,
                "potrf_info": info_big,
            },
            "info": info,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0       # fallback stated in B200_PROFILING.md


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
