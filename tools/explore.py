"""GPU exploration harness (not part of the product): FP64 pipe microbenchmarks and
per-stage timings under the two CTA shapes.  Usage: python tools/explore.py [what ...]"""
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

dev = torch.device("cuda:0")
lib.gp2d_dbg_fp64_mode.restype = C.c_int
lib.gp2d_dbg_fp64_mode.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]


def ev():
    return torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=3, warm=1):
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


def micro():
    out = torch.zeros(8, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    iters = 20000
    for ctas in (148, 296, 592):
        for mode in (0, 1, 2):
            t = timeit(lambda: lib.gp2d_dbg_fp64_mode(mode, iters, ctas, out.data_ptr(), st))
            dmma = ctas * 8 * iters * 16 * 512 if mode in (0, 2) else 0
            dfma = ctas * 256 * iters * (32 if mode == 1 else 16 if mode == 2 else 0) * 2
            print("micro ctas=%d (%d warps/SMSP) mode=%d: %.3f ms  DMMA %.2f TF/s  DFMA %.2f TF/s  sum %.2f" % (
                ctas, ctas * 8 // (148 * 4), mode, t * 1e3, dmma / t / 1e12, dfma / t / 1e12, (dmma + dfma) / t / 1e12))


def latency():
    out = torch.zeros(8, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    iters = 200000
    for mode, name in ((3, "DFMA"), (4, "DMMA")):
        t = timeit(lambda: lib.gp2d_dbg_fp64_mode(mode, iters, 1, out.data_ptr(), st))
        print("latency %s dependent chain: %.1f ns/op = %.1f cycles @1.965GHz" % (name, t / (16 * iters) * 1e9, t / (16 * iters) * 1.965e9))


def stages(N=2000, grid=(320, 320), theta=(1.3, 3.1, 0.2)):
    X, y = synthetic.drifter_snapshot(N, config_id=2)
    Xs = synthetic.prediction_grid(X, *grid)
    M = Xs.shape[0]
    n = 2 * N
    for nt in (288,):
        m = gp.HelmholtzGP(X, y, *theta, 0.05)
        Xsd = gp.as_dev(Xs)
        t_fit = timeit(lambda: m.fit_async())
        t_pred = timeit(lambda: m.predict(Xsd))
        t_grad = timeit(lambda: m.lml_and_grad())
        fl = float(n) * n * 2 * M + 2.0 * n * 2 * M
        print("stages nt=%d N=%d M=%d: fit %.3f ms  predict %.3f ms (%.2f TF/s)  lml+grad %.3f ms" % (
            nt, N, M, t_fit * 1e3, t_pred * 1e3, fl / t_pred / 1e12, t_grad * 1e3))
        m2 = gp.HelmholtzGP(X, y, 2.0, 2.0, 0.5, 0.05)
        t_pred2 = timeit(lambda: m2.predict(Xsd))
        print("   equal length scales (one exp): predict %.3f ms (%.2f TF/s)" % (t_pred2 * 1e3, fl / t_pred2 / 1e12))


def potrf_sizes():
    for nt in (288,):
        for n in (1024, 4096, 8192, 16384):
            g = torch.Generator(device=dev).manual_seed(n)
            B = torch.randn(n, 256, generator=g, dtype=torch.float64, device=dev)
            A = B @ B.t() / 256 + torch.eye(n, dtype=torch.float64, device=dev)
            nb_ws = lib.gp2d_potrf_workspace_bytes(n)
            ws = torch.empty(nb_ws, dtype=torch.uint8, device=dev)
            info = torch.zeros(1, dtype=torch.int32, device=dev)
            Aw = torch.empty_like(A)
            st = torch.cuda.current_stream().cuda_stream

            def run():
                Aw.copy_(A)
                lib.gp2d_potrf(Aw.data_ptr(), n, n, ws.data_ptr(), nb_ws, info.data_ptr(), st)
            t_copy = timeit(lambda: Aw.copy_(A))
            t = timeit(run) - t_copy
            print("potrf nt=%d n=%d: %.3f ms  %.2f TF/s (n^3/3)  info=%d" % (nt, n, t * 1e3, n ** 3 / 3 / t / 1e12, int(info.item())))
            del A, B, Aw, ws


def gemm():
    lib.gp2d_dbg_gemm.restype = C.c_int
    lib.gp2d_dbg_gemm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                  C.c_int, C.c_void_p]
    st = torch.cuda.current_stream().cuda_stream
    for nt in (288,):
        for (M, N, K) in [(8192, 8192, 8192), (4096, 4096, 4096), (16384, 16384, 128), (2048, 2048, 2048)]:
            A = torch.randn(max(M, K), max(M, K), dtype=torch.float64, device=dev)
            B = torch.randn(max(N, K), max(N, K), dtype=torch.float64, device=dev)
            Cm = torch.zeros(M, N, dtype=torch.float64, device=dev)
            for (a_mn, b_mn) in [(0, 0), (0, 1), (1, 1)]:
                for beta in (0.0, 1.0):
                    t = timeit(lambda: lib.gp2d_dbg_gemm(a_mn, b_mn, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                                         Cm.data_ptr(), Cm.stride(0), M, N, K, 1.0, beta, 0, 0, st))
                    print("gemm nt=%d %dx%dx%d a_mn=%d b_mn=%d beta=%g: %.3f ms %.2f TF/s" % (
                        nt, M, N, K, a_mn, b_mn, beta, t * 1e3, 2.0 * M * N * K / t / 1e12))
            del A, B, Cm


def rbf():
    """Scalar ARD-RBF path (krig.scikit_prior shape: (t, y, x) inputs, two components)."""
    rng = np.random.default_rng(1)
    for N, M in ((4000, 102400), (16384, 20000)):
        X = np.stack([rng.uniform(0, 24, N), rng.uniform(0, 40, N), rng.uniform(0, 40, N)], axis=1)
        y = np.sin(X[:, 1] / 5) + 0.05 * rng.normal(size=N)
        Xs = np.stack([np.full(M, 12.0), rng.uniform(0, 40, M), rng.uniform(0, 40, M)], axis=1)
        g = gp.ScalarGP(X, y, [0.05, 0.02], [[20.0, 6.0, 7.0], [3.0, 1.5, 2.0]], 0.0009, jitter=1e-10)
        Xsd = gp.as_dev(Xs)
        t_fit = timeit(lambda: g.fit_async())
        t_pred = timeit(lambda: g.predict(Xsd, include_noise=True))
        t_grad = timeit(lambda: g.lml_and_grad())
        fl = float(N) * N * M + 2.0 * N * M
        print("rbf N=%d M=%d: fit %.3f ms  predict %.3f ms (%.2f TF/s)  lml+grad %.3f ms" % (
            N, M, t_fit * 1e3, t_pred * 1e3, fl / t_pred / 1e12, t_grad * 1e3))


def split_calibration():
    """predict time vs forced nsplit: slope = phase-1 time per item (phi of predict_choose_split)."""
    lib.gp2d_dbg_set_predict_split.restype = C.c_int
    lib.gp2d_dbg_set_predict_split.argtypes = [C.c_int]
    X, y = synthetic.drifter_snapshot(2000, config_id=2)
    Xs = synthetic.prediction_grid(X, 320, 320)
    m = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05)
    m.fit()
    Xsd = gp.as_dev(Xs)
    rng = np.random.default_rng(1)
    N, M = 4000, 102400
    Xr = np.stack([rng.uniform(0, 24, N), rng.uniform(0, 40, N), rng.uniform(0, 40, N)], axis=1)
    yr = np.sin(Xr[:, 1] / 5) + 0.05 * rng.normal(size=N)
    Xrs = gp.as_dev(np.stack([np.full(M, 12.0), rng.uniform(0, 40, M), rng.uniform(0, 40, M)], axis=1))
    g = gp.ScalarGP(Xr, yr, [0.05, 0.02], [[20.0, 6.0, 7.0], [3.0, 1.5, 2.0]], 0.0009, jitter=1e-10)
    g.fit()
    for s in (1, 2, 4, 8):
        lib.gp2d_dbg_set_predict_split(s)
        m._pws = None
        g._pws = None
        th = timeit(lambda: m.predict(Xsd))
        tr = timeit(lambda: g.predict(Xrs))
        print("nsplit=%d: helmholtz predict %.3f ms   rbf predict %.3f ms" % (s, th * 1e3, tr * 1e3))
    lib.gp2d_dbg_set_predict_split(0)


def gemm_small():
    """128-tile warp-specialised kernel vs 64-tile kernel on the small GEMMs of the Cholesky recursion."""
    lib.gp2d_dbg_gemm.restype = C.c_int
    lib.gp2d_dbg_gemm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                  C.c_int, C.c_void_p]
    lib.gp2d_dbg_set_small_tile_threshold.restype = C.c_int
    lib.gp2d_dbg_set_small_tile_threshold.argtypes = [C.c_int]
    st = torch.cuda.current_stream().cuda_stream
    for n in (256, 512, 768, 1024, 1280, 1536, 2048, 3072):
        A = torch.randn(n, n, dtype=torch.float64, device=dev)
        B = torch.randn(n, n, dtype=torch.float64, device=dev)
        Cm = torch.zeros(n, n, dtype=torch.float64, device=dev)
        res = []
        for thr in (0, 1 << 30):
            lib.gp2d_dbg_set_small_tile_threshold(thr)
            for lower in (0, 1):
                t = timeit(lambda: lib.gp2d_dbg_gemm(0, 0, A.data_ptr(), n, B.data_ptr(), n, Cm.data_ptr(), n, n, n, n,
                                                     1.0, 0.0, lower, 0, st), reps=20, warm=3)
                res.append(t * 1e6)
        t128 = (n // 128) ** 2
        print("gemm n=%d (%d tiles / %d lower): ws128 %.1f us (lower %.1f)   tile64 %.1f us (lower %.1f)" % (
            n, t128, (n // 128) * (n // 128 + 1) // 2, res[0], res[1], res[2], res[3]))
    lib.gp2d_dbg_set_small_tile_threshold(-1)


def hsum(N=2000, grid=(320, 320)):
    """Term-sum family at the configs[1] size: fit, likelihood gradient and prediction against the
    Helmholtz family on the same data (Q = 2 terms, with and without the time factor)."""
    X, y = synthetic.drifter_snapshot(N, config_id=2)
    Xs = synthetic.prediction_grid(X, *grid)
    M, n = Xs.shape[0], 2 * N
    fl = float(n) * n * 2 * M
    rng = np.random.default_rng(0)
    X3 = np.ascontiguousarray(np.c_[rng.uniform(0, 6, N), X])
    Xs3 = np.ascontiguousarray(np.c_[np.full(M, 3.0), Xs])
    Xd, Xsd, X3d, Xs3d = (gp.as_dev(a) for a in (X, Xs, X3, Xs3))
    ref = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05)
    t_fit, t_pred, t_grad = timeit(ref.fit_async), timeit(lambda: ref.predict(Xsd)), timeit(ref.lml_and_grad)
    print("helmholtz        : fit %.3f ms, lml+grad %.3f ms, predict %.3f ms (%.2f TF/s)" % (t_fit * 1e3, t_grad * 1e3, t_pred * 1e3, fl / t_pred / 1e12))
    for name, pts, gpts, types, params in (
            ("hsum Q=2 (y,x)  ", Xd, Xsd, [0, 1], [[0.2, 1, 1.3, 1.3], [0.8, 1, 3.1, 3.1]]),
            ("hsum Q=2 (t,y,x)", X3d, Xs3d, [0, 1], [[0.2, 2.0, 1.3, 1.5], [0.8, 3.0, 3.1, 2.7]]),
            ("hsum Q=4 (t,y,x)", X3d, Xs3d, [0, 1, 0, 1], [[0.2, 2.0, 1.3, 1.5], [0.8, 3.0, 3.1, 2.7], [0.1, 1.0, 0.7, 0.6], [0.3, 5.0, 6.0, 5.0]]),
            ("hsum Q=8 (t,y,x)", X3d, Xs3d, [0, 1] * 4, [[0.2 + 0.05 * q, 2.0 + q, 1.3 + 0.2 * q, 1.5 + 0.1 * q] for q in range(8)])):
        g = gp.HelmholtzSumGP(pts, y, types, params, 0.05)
        t_fit, t_pred, t_grad = timeit(g.fit_async), timeit(lambda: g.predict(gpts)), timeit(g.lml_and_grad)
        print("%s: fit %.3f ms, lml+grad %.3f ms, predict %.3f ms (%.2f TF/s)" % (name, t_fit * 1e3, t_grad * 1e3, t_pred * 1e3, fl / t_pred / 1e12))


if __name__ == "__main__":
    what = sys.argv[1:] or ["micro", "stages", "potrf"]
    if "micro" in what:
        micro()
    if "latency" in what:
        latency()
    if "stages" in what:
        stages()
    if "potrf" in what:
        potrf_sizes()
    if "gemm" in what:
        gemm()
    if "split" in what:
        split_calibration()
    if "rbf" in what:
        rbf()
    if "hsum" in what:
        hsum()
    if "gemm_small" in what:
        gemm_small()
