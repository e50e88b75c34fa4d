"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden
vectors generated from the reference.  Run on the B200 box:  pytest tests -m gpu"""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():          # collected on the CPU box, run on the GPU box
    pytest.skip("no CUDA device", allow_module_level=True)

import gp2d_b200 as gp                      # noqa: E402
from gp2d_b200 import synthetic             # noqa: E402
from gp2d_b200._lib import lib              # noqa: E402
from oracle import gp_oracle as orc        # noqa: E402

DEV = torch.device("cuda:0")
THETAS = [(2.0, 2.0, 0.5), (1.3, 3.1, 0.2), (0.6, 0.6, 1.0), (0.7, 1.9, 0.0)]


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)


def stream():
    return torch.cuda.current_stream().cuda_stream


# ---- bring-up hooks ---------------------------------------------------------------------
lib.gp2d_dbg_gemm.restype = C.c_int
lib.gp2d_dbg_gemm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                              C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                              C.c_int, C.c_void_p]
lib.gp2d_dbg_potri.restype = C.c_int
lib.gp2d_dbg_potri.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p]


lib.gp2d_dbg_set_small_tile_threshold.restype = C.c_int
lib.gp2d_dbg_set_small_tile_threshold.argtypes = [C.c_int]


@pytest.fixture(params=[0, 1 << 30], ids=["tile128", "tile64"])
def tile_variant(request):
    """Force the 128x128-tile or the 64x64-tile GEMM kernel for every launch in the test."""
    lib.gp2d_dbg_set_small_tile_threshold(request.param)
    yield request.param
    lib.gp2d_dbg_set_small_tile_threshold(-1)


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_dgemm_layouts(a_mn, b_mn, tile_variant):
    g = torch.Generator(device="cpu").manual_seed(1)
    M, N, K = 256, 384, 160
    A = torch.randn(M, K, generator=g, dtype=torch.float64).to(DEV)
    B = torch.randn(K, N, generator=g, dtype=torch.float64).to(DEV)
    C0 = torch.randn(M, N, generator=g, dtype=torch.float64).to(DEV)
    Ast = A.t().contiguous() if a_mn else A.contiguous()          # MN-major: stored [K][M]
    Bst = B.contiguous() if b_mn else B.t().contiguous()          # K-major: stored [N][K]
    Cc = C0.clone()
    rc = lib.gp2d_dbg_gemm(a_mn, b_mn, Ast.data_ptr(), Ast.stride(0), Bst.data_ptr(), Bst.stride(0),
                           Cc.data_ptr(), Cc.stride(0), M, N, K, -1.5, 0.75, 0, 0, stream())
    assert rc == 0, lib.gp2d_error_string(rc)
    ref = -1.5 * (A @ B) + 0.75 * C0
    torch.testing.assert_close(Cc, ref, rtol=1e-12, atol=1e-11)


def test_dgemm_triangular_clipping(tile_variant):
    g = torch.Generator(device="cpu").manual_seed(2)
    n = 384
    Zl = torch.tril(torch.randn(n, n, generator=g, dtype=torch.float64)).to(DEV)
    ti = torch.arange(n, device=DEV) // 128
    tile_upper = (ti[None, :] > ti[:, None]).to(torch.float64)
    Zgarb = Zl + 7.0 * tile_upper              # never-read tiles hold garbage
    D = torch.randn(n, n, generator=g, dtype=torch.float64).to(DEV)
    out = torch.zeros(n, n, dtype=torch.float64, device=DEV)
    # KR_LE_N (2): C = D * Z^T with Z lower (upper off-diagonal tiles hold garbage)
    rc = lib.gp2d_dbg_gemm(0, 0, D.data_ptr(), n, Zgarb.data_ptr(), n, out.data_ptr(), n, n, n, n, 1.0, 0.0, 0, 2, stream())
    assert rc == 0
    torch.testing.assert_close(out, D @ Zl.t(), rtol=1e-12, atol=1e-11)
    # KR_GE_N (4): C = D * Z (B MN-major)
    rc = lib.gp2d_dbg_gemm(0, 1, D.data_ptr(), n, Zgarb.data_ptr(), n, out.data_ptr(), n, n, n, n, 1.0, 0.0, 0, 4, stream())
    assert rc == 0
    torch.testing.assert_close(out, D @ Zl, rtol=1e-12, atol=1e-11)
    # KR_LE_M (1): C = -Z * D (A lower, B MN-major)
    rc = lib.gp2d_dbg_gemm(0, 1, Zgarb.data_ptr(), n, D.data_ptr(), n, out.data_ptr(), n, n, n, n, -1.0, 0.0, 0, 1, stream())
    assert rc == 0
    torch.testing.assert_close(out, -(Zl @ D), rtol=1e-12, atol=1e-11)
    # KR_GE_M (8) + lower_out: lower tiles of Z^T Z
    out.zero_()
    rc = lib.gp2d_dbg_gemm(1, 1, Zgarb.data_ptr(), n, Zgarb.data_ptr(), n, out.data_ptr(), n, n, n, n, 1.0, 0.0, 1, 8, stream())
    assert rc == 0
    ref = Zl.t() @ Zl
    tile_lower = (torch.arange(n, device=DEV)[None, :] // 128 <= torch.arange(n, device=DEV)[:, None] // 128)
    # the contract: the lower triangle is written; nothing above the diagonal 128-tiles is touched
    # (the 64-tile variant also leaves the upper-right quadrant of diagonal 128-tiles alone)
    torch.testing.assert_close(torch.tril(out), torch.tril(ref), rtol=1e-12, atol=1e-11)
    assert float((out * ~tile_lower).abs().max()) == 0.0


@pytest.mark.parametrize("n", [128, 256, 384, 1024])
def test_potri_recursion(n, tile_variant):
    g = torch.Generator(device="cpu").manual_seed(n)
    Bm = torch.randn(n, n, generator=g, dtype=torch.float64)
    A = (Bm @ Bm.t() / n + torch.eye(n, dtype=torch.float64)).to(DEV)
    Lref = torch.linalg.cholesky(A)
    for need_inv, keep_L in [(1, 0), (1, 1), (0, 1)]:
        Aw = A.clone()
        Z = torch.zeros(n, n, dtype=torch.float64, device=DEV)
        W = torch.zeros(n * n // 4 + 64, dtype=torch.float64, device=DEV)
        logd = torch.zeros(n, dtype=torch.float64, device=DEV)
        info = torch.zeros(1, dtype=torch.int32, device=DEV)
        rc = lib.gp2d_dbg_potri(Aw.data_ptr(), n, Z.data_ptr(), logd.data_ptr(), info.data_ptr(), need_inv,
                                keep_L, W.data_ptr(), stream())
        assert rc == 0, lib.gp2d_error_string(rc)
        assert int(info.item()) == 0
        torch.testing.assert_close(logd, torch.log(torch.diagonal(Lref)), rtol=1e-11, atol=1e-12)
        if keep_L:
            torch.testing.assert_close(torch.tril(Aw), Lref, rtol=1e-10, atol=1e-11)
        if need_inv:
            Zi = torch.linalg.inv(Lref)
            torch.testing.assert_close(torch.tril(Z), Zi, rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 300, 1000])
def test_potrf_public(n):
    g = torch.Generator(device="cpu").manual_seed(100 + n)
    Bm = torch.randn(n, n, generator=g, dtype=torch.float64)
    A = (Bm @ Bm.t() / n + torch.eye(n, dtype=torch.float64)).to(DEV)
    L, info = gp.potrf(A)
    assert info == 0
    torch.testing.assert_close(L, torch.linalg.cholesky(A), rtol=1e-10, atol=1e-11)


def test_potrf_not_positive_definite():
    n = 200
    A = torch.eye(n, dtype=torch.float64, device=DEV)
    A[150, 150] = -1.0
    _, info = gp.potrf(A)
    assert info == 151


# ---- kernel build -------------------------------------------------------------------------
def test_kernel_build_golden(golden_dir):
    ks = np.load(os.path.join(golden_dir, "kernel_small.npz"))
    X, X2 = ks["X"], ks["X2"]
    for t, (ldf, lcf, r) in enumerate(ks["thetas"]):
        K = gp.kernel_K(X, None, ldf, lcf, r).cpu().numpy()
        np.testing.assert_allclose(K, ks["ref_K_class_sym_%d" % t], rtol=0, atol=2e-14)
        K = gp.kernel_K(X, X2, ldf, lcf, r).cpu().numpy()
        np.testing.assert_allclose(K, ks["ref_K_class_x_%d" % t], rtol=0, atol=2e-14)
        # compute_Ks orientation: K(X*, X)
        K = gp.kernel_K(X2, X, ldf, lcf, r).cpu().numpy()
        np.testing.assert_allclose(K, ks["ref_Ks_loops_%d" % t], rtol=0, atol=2e-14)
        d = gp.kernel_Kdiag(X2.shape[0], ldf, lcf, r).cpu().numpy()
        np.testing.assert_allclose(d, ks["ref_Kdiag_%d" % t], rtol=1e-15)


@pytest.mark.parametrize("N,M", [(1, 1), (3, 7), (33, 64), (130, 129), (257, 1000)])
def test_kernel_build_ragged_vs_oracle(N, M):
    rng = np.random.default_rng(N * 1000 + M)
    X = rng.uniform(0, 20, size=(N, 2))
    X2 = rng.uniform(0, 20, size=(M, 2))
    for (ldf, lcf, r) in THETAS[:2]:
        K = gp.kernel_K(X, X2, ldf, lcf, r).cpu().numpy()
        np.testing.assert_allclose(K, orc.helmholtz_K(X, X2, ldf, lcf, r), rtol=0, atol=1e-14)
    Ks = gp.kernel_K(X, None, 2.0, 2.0, 0.5, diag_add=0.05).cpu().numpy()
    ref = orc.helmholtz_K(X, None, 2.0, 2.0, 0.5) + 0.05 * np.eye(2 * N)
    np.testing.assert_allclose(Ks, ref, rtol=0, atol=1e-14)
    # strided output (ld > 2M)
    out = torch.full((2 * N, 2 * M + 3), -7.0, dtype=torch.float64, device=DEV)
    gp.kernel_K(X, X2, 1.3, 3.1, 0.2, out=out[:, :2 * M])
    np.testing.assert_allclose(out[:, :2 * M].cpu().numpy(), orc.helmholtz_K(X, X2, 1.3, 3.1, 0.2), atol=1e-14)
    assert float(out[:, 2 * M:].min()) == -7.0 and float(out[:, 2 * M:].max()) == -7.0


def test_kernel_far_points_underflow():
    X = np.array([[0.0, 0.0], [1e4, -1e4], [3.0, 4.0]])
    K = gp.kernel_K(X, None, 0.5, 0.7, 0.3).cpu().numpy()
    np.testing.assert_allclose(K, orc.helmholtz_K(X, None, 0.5, 0.7, 0.3), rtol=0, atol=1e-15)
    assert np.isfinite(K).all()


def test_kernel_grad_sums_golden(golden_dir):
    ks = np.load(os.path.join(golden_dir, "kernel_small.npz"))
    X, X2 = ks["X"], ks["X2"]
    for t, (ldf, lcf, r) in enumerate(ks["thetas"]):
        g = gp.kernel_grad_sums(ks["W_sym"], X, None, ldf, lcf, r, reference_compat=True).cpu().numpy()
        np.testing.assert_allclose(g, ks["ref_grad_compat_sym_%d" % t], rtol=1e-11, atol=1e-12)
        g = gp.kernel_grad_sums(ks["W_x"], X, X2, ldf, lcf, r, reference_compat=True).cpu().numpy()
        np.testing.assert_allclose(g, ks["ref_grad_compat_x_%d" % t], rtol=1e-11, atol=1e-12)
        g = gp.kernel_grad_sums(ks["W_x"], X, X2, ldf, lcf, r).cpu().numpy()
        np.testing.assert_allclose(g, orc.kernel_grad_sums(ks["W_x"], X, X2, ldf, lcf, r), rtol=1e-11, atol=1e-12)


# ---- fit / predict / likelihood -------------------------------------------------------------
@pytest.mark.parametrize("ts", [0, 100])
def test_simlaser_fit_predict(golden_dir, ts):
    """Config 1: simulTracks.pkl snapshot through GP_laser.simLaser's defaults, against the
    reference's own numpy pipeline (tests/golden/make_golden.py).  Tolerances from
    BASELINE.json: 1e-8 relative on mean/variance, 1e-6 on the log-likelihood."""
    g = np.load(os.path.join(golden_dir, "simlaser_ts%d.npz" % ts))
    X, y, Xs = g["X"], g["y"], g["Xs"]
    ldf, lcf, r = g["theta"]
    m = gp.HelmholtzGP(X, y, ldf, lcf, r, float(g["noise"]))
    lml = m.fit()
    mean, var = m.predict(Xs)
    mean, var = mean.cpu().numpy(), var.cpu().numpy()
    scale = np.abs(g["ref_mean"]).max()
    np.testing.assert_allclose(mean, g["ref_mean"], rtol=1e-8, atol=1e-8 * scale)
    np.testing.assert_allclose(var, g["ref_var"], rtol=1e-8)
    assert abs(lml - float(g["derived_lml"])) <= 1e-6 * abs(float(g["derived_lml"]))
    f = orc.fit(X, y, ldf, lcf, r, float(g["noise"]))
    np.testing.assert_allclose(m.alpha().cpu().numpy(), f["alpha"], rtol=1e-8, atol=1e-9 * np.abs(f["alpha"]).max())
    # GPy convention: noise variance included
    _, var_n = m.predict(Xs, include_noise=True)
    np.testing.assert_allclose(var_n.cpu().numpy(), g["ref_var"] + float(g["noise"]), rtol=1e-8)


@pytest.mark.parametrize("N,M,theta", [(1, 3, THETAS[0]), (7, 1, THETAS[1]), (64, 64, THETAS[0]),
                                       (16, 70, THETAS[1]), (40, 65, THETAS[2]), (48, 129, THETAS[1]), (80, 64, THETAS[3]),
                                       (1040, 300, THETAS[1]), (1072, 90, THETAS[0]),
                                       (65, 130, THETAS[1]), (300, 517, THETAS[2]), (513, 200, THETAS[3])])
def test_fit_predict_vs_oracle(N, M, theta):
    X, y = synthetic.drifter_snapshot(N, config_id=9, seed_offset=N)
    Xs = synthetic.prediction_grid(X, M, 1)
    ldf, lcf, r = theta
    noise = 0.05
    m = gp.HelmholtzGP(X, y, ldf, lcf, r, noise)
    lml = m.fit()
    mean, var = m.predict(Xs)
    f = orc.fit(X, y, ldf, lcf, r, noise)
    mo, vo = orc.predict(X, f, ldf, lcf, r, Xs)
    np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-8, atol=1e-9 * max(np.abs(mo).max(), 1e-3))
    np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-8, atol=1e-12)
    assert abs(lml - f["lml"]) <= 1e-6 * max(abs(f["lml"]), 1.0)


def test_lml_grad_vs_oracle():
    X, y = synthetic.drifter_snapshot(200, config_id=4)
    for theta in [(2.0, 2.0, 0.5), (1.3, 3.1, 0.2)]:
        m = gp.HelmholtzGP(X, y, *theta, 0.05, jitter=1e-8)
        for compat in (False, True):
            lml, grad = m.lml_and_grad(reference_compat=compat)
            lo, go = orc.lml_and_grad(X, y, *theta, 0.05, jitter=1e-8, reference_compat=compat)
            assert abs(lml - lo) <= 1e-6 * abs(lo)
            np.testing.assert_allclose(grad, go, rtol=1e-6, atol=1e-7)
        # a valid fit state is left behind
        mean, _ = m.predict(X[:5])
        f = orc.fit(X, y, *theta, 0.05, jitter=1e-8)
        mo, _ = orc.predict(X, f, *theta, X[:5])
        np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("N", [100, 700, 1100])
def test_predict_row_block_split_invariance(N):
    """Small grids spread the row blocks of a column tile over several CTAs (nsplit up to 8); the
    grouping and summation order are fixed, so any grid size gives the same bits, and the oracle
    agrees."""
    X, y = synthetic.drifter_snapshot(N, config_id=8, seed_offset=N)
    big = synthetic.prediction_grid(X, 160, 120)                 # 19 200 points: 300 tiles
    m = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05)
    m.fit()
    mb, vb = m.predict(big)
    Mb = big.shape[0]
    for M in (1, 63, 64, 65, 640, 2601, 9472):
        mm, vv = m.predict(big[:M])
        assert torch.equal(mm, torch.cat([mb[:M], mb[Mb:Mb + M]])), M
        assert torch.equal(vv, torch.cat([vb[:M], vb[Mb:Mb + M]])), M
    # a scratch buffer too small for the split falls back to one CTA per tile: same bits
    from gp2d_b200.engine import _ptr, _stream
    M = 640
    Xsd = gp.as_dev(big[:M])
    small = torch.empty(8 * 128 * (((2 * N + 127) // 128) * 128) + 20480, dtype=torch.uint8, device=DEV)   # exactly one K* panel
    mean = torch.empty(2 * M, dtype=torch.float64, device=DEV)
    var = torch.empty(2 * M, dtype=torch.float64, device=DEV)
    rc = lib.gp2d_predict(_ptr(m.ws), N, 1.3, 3.1, 0.2, _ptr(Xsd), M, M, 0.0, _ptr(mean), _ptr(var), _ptr(small),
                          small.numel(), _stream())
    assert rc == 0
    assert torch.equal(mean, torch.cat([mb[:M], mb[Mb:Mb + M]])) and torch.equal(var, torch.cat([vb[:M], vb[Mb:Mb + M]]))
    f = orc.fit(X, y, 1.3, 3.1, 0.2, 0.05)
    mo, vo = orc.predict(X, f, 1.3, 3.1, 0.2, big[:2601])
    mm, vv = m.predict(big[:2601])
    np.testing.assert_allclose(mm.cpu().numpy(), mo, rtol=1e-8, atol=1e-9 * np.abs(mo).max())
    np.testing.assert_allclose(vv.cpu().numpy(), vo, rtol=1e-8, atol=1e-12)


def test_predict_state_ships_to_another_workspace():
    """What dist.broadcast_fit relies on: the byte range gp2d_fit_predict_state names is all that
    gp2d_predict reads, so a copy of it in a fresh workspace predicts bit-identically."""
    X, y = synthetic.drifter_snapshot(300, config_id=7)
    Xs = synthetic.prediction_grid(X, 23, 11)
    a = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05)
    a.fit()
    m0, v0 = a.predict(Xs)
    b = gp.HelmholtzGP(X * 0 + 5.0, y * 0, 1.3, 3.1, 0.2, 0.05)      # different data, never fitted
    b.ws.fill_(255)
    b.predict_state().copy_(a.predict_state())
    b.fitted = True
    m1, v1 = b.predict(Xs)
    assert torch.equal(m0, m1) and torch.equal(v0, v1)
    assert a.predict_state().numel() < a.ws.numel() // 2


def test_side_stream_overlap_changes_nothing():
    """The inverse-update GEMMs run on side streams beside the Cholesky chain; with the overlap
    switched off the same kernels run on one stream: identical bits either way."""
    lib.gp2d_dbg_set_potri_overlap.restype = C.c_int
    lib.gp2d_dbg_set_potri_overlap.argtypes = [C.c_int]
    X, y = synthetic.drifter_snapshot(1100, config_id=6)
    Xs = synthetic.prediction_grid(X, 30, 30)
    res = []
    for on in (1, 0, 1):
        lib.gp2d_dbg_set_potri_overlap(on)
        m = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05)
        lml, grad = m.lml_and_grad()
        al = m.alpha()
        mean, var = m.predict(Xs)
        res.append((lml, grad, al, mean, var))
    lib.gp2d_dbg_set_potri_overlap(1)
    for r in res[1:]:
        assert r[0] == res[0][0] and np.array_equal(r[1], res[0][1])
        assert torch.equal(r[2], res[0][2]) and torch.equal(r[3], res[0][3]) and torch.equal(r[4], res[0][4])
    f = orc.fit(X, y, 1.3, 3.1, 0.2, 0.05)
    np.testing.assert_allclose(res[0][2].cpu().numpy(), f["alpha"], rtol=1e-7, atol=1e-9 * np.abs(f["alpha"]).max())


def test_not_positive_definite_raises():
    X = np.zeros((40, 2))                     # coincident points, no noise -> singular
    y = np.ones(80)
    m = gp.HelmholtzGP(X, y, 1.0, 1.0, 0.5, 0.0)
    with pytest.raises(np.linalg.LinAlgError):
        m.fit()


def test_host_pointer_entry_point(golden_dir):
    g = np.load(os.path.join(golden_dir, "simlaser_ts0.npz"))
    ldf, lcf, r = g["theta"]
    mean, var, lml = gp.fit_predict_host(g["X"], g["y"], ldf, lcf, r, float(g["noise"]), g["Xs"])
    np.testing.assert_allclose(mean, g["ref_mean"], rtol=1e-8, atol=1e-8 * np.abs(g["ref_mean"]).max())
    np.testing.assert_allclose(var, g["ref_var"], rtol=1e-8)
    assert abs(lml - float(g["derived_lml"])) < 1e-6 * abs(float(g["derived_lml"]))


def test_bad_arguments():
    X = dev(np.zeros((4, 2)))
    out = torch.zeros(8, 8, dtype=torch.float64, device=DEV)
    assert lib.gp2d_kernel_build(X.data_ptr(), 4, None, 4, -1.0, 1.0, 0.5, 0.0, out.data_ptr(), 8, None) == -5
    assert lib.gp2d_kernel_build(None, 4, None, 4, 1.0, 1.0, 0.5, 0.0, out.data_ptr(), 8, None) == -1
    assert lib.gp2d_kernel_build(X.data_ptr(), 4, None, 4, 1.0, 1.0, 0.5, 0.0, out.data_ptr(), 7, None) == -10
    assert lib.gp2d_fit(X.data_ptr(), 4, out.data_ptr(), 1.0, 1.0, 0.5, 0.1, 0.0, out.data_ptr(), 16, None, None,
                        None, None) == -10
    # sizes beyond what one GPU can factorise are refused, not attempted
    assert lib.gp2d_fit_workspace_bytes(70000) == 0
    assert lib.gp2d_fit(X.data_ptr(), 70000, out.data_ptr(), 1.0, 1.0, 0.5, 0.1, 0.0, out.data_ptr(), 16, None, None,
                        None, None) == -2
    assert lib.gp2d_error_string(-2).decode() == "invalid argument 2"
    assert b"CUDA error" in lib.gp2d_error_string(-1001)


# ---- full-size configuration: size-independent properties + sampled oracle -------------------
def test_config2_full_size_properties():
    """BASELINE.json configs[1]: N=2000 (4k x 4k covariance), 102 400-point grid."""
    N, nx, ny = 2000, 320, 320
    X, y = synthetic.drifter_snapshot(N, config_id=2)
    Xs = synthetic.prediction_grid(X, nx, ny)
    theta, noise = (1.3, 3.1, 0.2), 0.05
    m = gp.HelmholtzGP(X, y, *theta, noise)
    lml = m.fit()
    mean, var = m.predict(Xs)
    M = Xs.shape[0]
    kss = orc.helmholtz_Kdiag(1, *theta)[0]
    assert float(var.min()) >= 0.0 and float(var.max()) <= kss * (1 + 1e-12)
    # sampled oracle (the CPU fit at this size takes seconds)
    f = orc.fit(X, y, *theta, noise)
    assert abs(lml - f["lml"]) <= 1e-6 * abs(f["lml"])
    idx = np.random.default_rng(0).choice(M, 300, replace=False)
    mo, vo = orc.predict(X, f, *theta, Xs[idx])
    mg = np.concatenate([mean[idx].cpu().numpy(), mean[M + idx].cpu().numpy()])
    vg = np.concatenate([var[idx].cpu().numpy(), var[M + idx].cpu().numpy()])
    np.testing.assert_allclose(mg, mo, rtol=1e-8, atol=1e-9 * np.abs(mo).max())
    np.testing.assert_allclose(vg, vo, rtol=1e-8, atol=1e-12)
    # shard invariance: any partition of the grid gives bit-identical results
    cut = 40001
    m1, v1 = m.predict(Xs[:cut])
    m2, v2 = m.predict(Xs[cut:])
    assert torch.equal(torch.cat([m1[:cut], m2[:M - cut], m1[cut:], m2[M - cut:]]), mean)
    assert torch.equal(torch.cat([v1[:cut], v2[:M - cut], v1[cut:], v2[M - cut:]]), var)
    # linearity of the mean in the observations
    y2 = np.roll(y, 17)
    ma = gp.HelmholtzGP(X, y2, *theta, noise); ma.fit()
    mb = gp.HelmholtzGP(X, y + 2.0 * y2, *theta, noise); mb.fit()
    pa, _ = ma.predict(Xs[:5000]); pb, _ = mb.predict(Xs[:5000]); p0, _ = m.predict(Xs[:5000])
    torch.testing.assert_close(pb, p0 + 2.0 * pa, rtol=1e-8, atol=1e-10)
    # GP identity at the observation sites: K alpha = y - noise * alpha
    pm, _ = m.predict(X)
    torch.testing.assert_close(pm + noise * m.alpha(), m.y, rtol=1e-9, atol=1e-10)


# ---- guard bands: no kernel writes outside the buffers the C ABI was given ----------------------
# (compute-sanitizer is closed on this GPU pool, so out-of-bounds writes are hunted this way)
GUARD = 8192


def _guarded(nbytes):
    """uint8 tensor of nbytes with GUARD bytes of 0xA5 on both sides; returns (whole, interior view)."""
    nbytes = (int(nbytes) + 255) // 256 * 256
    whole = torch.full((nbytes + 2 * GUARD,), 0xA5, dtype=torch.uint8, device=DEV)
    return whole, whole[GUARD:GUARD + nbytes]


def _intact(whole):
    return bool((whole[:GUARD] == 0xA5).all()) and bool((whole[-GUARD:] == 0xA5).all())


@pytest.mark.parametrize("N,M", [(1, 1), (63, 65), (200, 129), (321, 1000), (1100, 5000)])
def test_no_writes_outside_buffers_helmholtz(N, M):
    from gp2d_b200.engine import _stream
    X, y = synthetic.drifter_snapshot(N, config_id=11, seed_offset=N)
    Xs = synthetic.prediction_grid(X, M, 1)
    Xd, yd, Xsd = dev(X), dev(y), dev(Xs)
    theta = (1.3, 3.1, 0.2)
    wsW, ws = _guarded(lib.gp2d_fit_workspace_bytes(N))
    pwW, pws = _guarded(lib.gp2d_predict_workspace_bytes(N, M))
    outW, out = _guarded(8 * (4 * M + 2 * N + 8 + 2))
    o = out.view(torch.float64)
    mean, var, alpha, scal = o[:2 * M], o[2 * M:4 * M], o[4 * M:4 * M + 2 * N], o[4 * M + 2 * N:4 * M + 2 * N + 8]
    info = torch.zeros(1, dtype=torch.int32, device=DEV)
    assert lib.gp2d_lml_grad(Xd.data_ptr(), N, yd.data_ptr(), *theta, 0.05, 1e-8, 0, ws.data_ptr(), ws.numel(),
                             scal.data_ptr(), info.data_ptr(), _stream()) == 0
    assert lib.gp2d_fit(Xd.data_ptr(), N, yd.data_ptr(), *theta, 0.05, 0.0, ws.data_ptr(), ws.numel(), alpha.data_ptr(),
                        scal.data_ptr(), info.data_ptr(), _stream()) == 0
    assert lib.gp2d_predict(ws.data_ptr(), N, *theta, Xsd.data_ptr(), M, M, 0.0, mean.data_ptr(), var.data_ptr(),
                            pws.data_ptr(), pws.numel(), _stream()) == 0
    # kernel build into a matrix with a wider leading dimension: the padding columns stay untouched
    ld = 2 * M + 6
    KW, Kb = _guarded(8 * 2 * N * ld)
    assert lib.gp2d_kernel_build(Xd.data_ptr(), N, Xsd.data_ptr(), M, *theta, 0.0, Kb.data_ptr(), ld, _stream()) == 0
    torch.cuda.synchronize()
    assert _intact(wsW) and _intact(pwW) and _intact(outW) and _intact(KW)
    Kv = Kb.view(torch.float64)[:2 * N * ld].view(2 * N, ld)
    assert bool((Kv[:, 2 * M:].contiguous().view(torch.uint8) == 0xA5).all())
    f = orc.fit(X, y, *theta, 0.05)
    mo, vo = orc.predict(X, f, *theta, Xs)
    np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-8, atol=1e-9 * max(np.abs(mo).max(), 1e-3))
    np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("N,M,D,Q", [(1, 1, 1, 1), (127, 129, 3, 2), (500, 3000, 3, 2), (1025, 700, 4, 4)])
def test_no_writes_outside_buffers_rbf(N, M, D, Q):
    from gp2d_b200.engine import _stream
    rng = np.random.default_rng(N)
    X, Xs = rng.uniform(0, 12, (N, D)), rng.uniform(0, 12, (M, D))
    y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=N)
    var_h = np.ascontiguousarray(rng.uniform(0.2, 1.5, Q))
    ls_h = np.ascontiguousarray(rng.uniform(1.0, 5.0, (Q, D)))
    Xd, yd, Xsd = dev(X), dev(y), dev(Xs)
    wsW, ws = _guarded(lib.gp2d_rbf_fit_workspace_bytes(N, D))
    pwW, pws = _guarded(lib.gp2d_rbf_predict_workspace_bytes(N, M))
    ng = 2 + Q * (1 + D)
    outW, out = _guarded(8 * (2 * M + N + ng))
    o = out.view(torch.float64)
    mean, var, alpha, scal = o[:M], o[M:2 * M], o[2 * M:2 * M + N], o[2 * M + N:2 * M + N + ng]
    info = torch.zeros(1, dtype=torch.int32, device=DEV)
    args = (D, )
    assert lib.gp2d_rbf_lml_grad(Xd.data_ptr(), N, D, yd.data_ptr(), Q, var_h.ctypes.data, ls_h.ctypes.data, 0.01, 1e-8,
                                 ws.data_ptr(), ws.numel(), scal.data_ptr(), info.data_ptr(), _stream()) == 0
    assert lib.gp2d_rbf_fit(Xd.data_ptr(), N, D, yd.data_ptr(), Q, var_h.ctypes.data, ls_h.ctypes.data, 0.01, 1e-8,
                            ws.data_ptr(), ws.numel(), alpha.data_ptr(), scal.data_ptr(), info.data_ptr(), _stream()) == 0
    assert lib.gp2d_rbf_predict(ws.data_ptr(), N, D, Q, var_h.ctypes.data, ls_h.ctypes.data, Xsd.data_ptr(), M, 0.01,
                                mean.data_ptr(), var.data_ptr(), pws.data_ptr(), pws.numel(), _stream()) == 0
    ld = M + 3
    KW, Kb = _guarded(8 * N * ld)
    assert lib.gp2d_rbf_kernel_build(Xd.data_ptr(), N, Xsd.data_ptr(), M, D, Q, var_h.ctypes.data, ls_h.ctypes.data, 0.0,
                                     Kb.data_ptr(), ld, _stream()) == 0
    torch.cuda.synchronize()
    assert _intact(wsW) and _intact(pwW) and _intact(outW) and _intact(KW)
    Kv = Kb.view(torch.float64)[:N * ld].view(N, ld)
    assert bool((Kv[:, M:].contiguous().view(torch.uint8) == 0xA5).all())
    np.testing.assert_allclose(Kv[:, :M].cpu().numpy(), orc.rbf_sum_K(X, Xs, var_h, ls_h), rtol=1e-12, atol=1e-15)
    f = orc.rbf_fit(X, y, var_h, ls_h, 0.01, jitter=1e-8)
    mo, vo = orc.rbf_predict(X, f, var_h, ls_h, Xs, var_add=0.01)
    np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-8, atol=1e-8 * max(np.abs(mo).max(), 1e-3))
    np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("N,M,ldx,Q", [(1, 1, 3, 1), (127, 129, 2, 2), (500, 3000, 3, 4), (513, 700, 3, 8)])
def test_no_writes_outside_buffers_hsum(N, M, ldx, Q):
    from gp2d_b200.engine import _stream
    rng = np.random.default_rng(N)
    X, Xs = rng.uniform(0, 12, (N, ldx)), rng.uniform(0, 12, (M, ldx))
    y = np.concatenate([np.sin(X[:, -1]), np.cos(X[:, -2])]) + 0.1 * rng.normal(size=2 * N)
    ty_h = np.ascontiguousarray(rng.integers(0, 2, Q).astype(np.int32))
    pr_h = np.ascontiguousarray(np.c_[rng.uniform(0.2, 1.5, Q), rng.uniform(1.0, 5.0, (Q, 3))])
    Xd, yd, Xsd = dev(X), dev(y), dev(Xs)
    wsW, ws = _guarded(lib.gp2d_hsum_fit_workspace_bytes(N, ldx, Q))
    pwW, pws = _guarded(lib.gp2d_predict_workspace_bytes(N, M))
    ng = 2 + 4 * Q
    outW, out = _guarded(8 * (4 * M + 2 * N + ng + 1 + 4 * Q))
    o = out.view(torch.float64)
    mean, var, alpha = o[:2 * M], o[2 * M:4 * M], o[4 * M:4 * M + 2 * N]
    scal = o[4 * M + 2 * N:4 * M + 2 * N + ng]
    lml1 = o[4 * M + 2 * N + ng:4 * M + 2 * N + ng + 1]
    gsum = o[4 * M + 2 * N + ng + 1:4 * M + 2 * N + ng + 1 + 4 * Q]
    info = torch.zeros(1, dtype=torch.int32, device=DEV)
    tp, pp = ty_h.ctypes.data, pr_h.ctypes.data
    assert lib.gp2d_hsum_lml_grad(Xd.data_ptr(), N, ldx, yd.data_ptr(), Q, tp, pp, 0.01, 1e-8, ws.data_ptr(), ws.numel(),
                                  scal.data_ptr(), info.data_ptr(), _stream()) == 0
    assert lib.gp2d_hsum_fit(Xd.data_ptr(), N, ldx, yd.data_ptr(), Q, tp, pp, 0.01, 1e-8, ws.data_ptr(), ws.numel(),
                             alpha.data_ptr(), lml1.data_ptr(), info.data_ptr(), _stream()) == 0
    assert lib.gp2d_hsum_predict(ws.data_ptr(), N, ldx, Q, tp, pp, Xsd.data_ptr(), M, M, 0.01, mean.data_ptr(), var.data_ptr(),
                                 pws.data_ptr(), pws.numel(), _stream()) == 0
    ld = 2 * M + 4
    KW, Kb = _guarded(8 * 2 * N * ld)
    assert lib.gp2d_hsum_kernel_build(Xd.data_ptr(), N, Xsd.data_ptr(), M, ldx, Q, tp, pp, 0.0, Kb.data_ptr(), ld, _stream()) == 0
    Wd = dev(rng.normal(size=(2 * N, 2 * M)))
    gW, gws = _guarded(lib.gp2d_hsum_kernel_grad_workspace_bytes(N, M, Q))
    assert lib.gp2d_hsum_kernel_grad(Xd.data_ptr(), N, Xsd.data_ptr(), M, ldx, Q, tp, pp, Wd.data_ptr(), 2 * M, gws.data_ptr(),
                                     gws.numel(), gsum.data_ptr(), _stream()) == 0
    torch.cuda.synchronize()
    assert _intact(wsW) and _intact(pwW) and _intact(outW) and _intact(KW) and _intact(gW)
    Kv = Kb.view(torch.float64)[:2 * N * ld].view(2 * N, ld)
    assert bool((Kv[:, 2 * M:].contiguous().view(torch.uint8) == 0xA5).all())
    np.testing.assert_allclose(Kv[:, :2 * M].cpu().numpy(), orc.hsum_K(X, Xs, ty_h, pr_h), rtol=0, atol=1e-13)
    f = orc.hsum_fit(X, y, ty_h, pr_h, 0.01, jitter=1e-8)
    mo, vo = orc.hsum_predict(X, f, ty_h, pr_h, Xs, var_add=0.01)
    np.testing.assert_allclose(mean.cpu().numpy(), mo, rtol=1e-8, atol=1e-8 * max(np.abs(mo).max(), 1e-3))
    np.testing.assert_allclose(var.cpu().numpy(), vo, rtol=1e-8, atol=1e-12)
    assert abs(float(lml1[0]) - f["lml"]) <= 1e-6 * abs(f["lml"]) and float(scal[0]) == float(lml1[0])
    np.testing.assert_allclose(gsum.cpu().numpy().reshape(Q, 4), orc.hsum_kernel_grad_sums(Wd.cpu().numpy(), X, Xs, ty_h, pr_h),
                               rtol=1e-9, atol=1e-9 * np.sqrt(N * M))


def test_config3_full_size_properties():
    """BASELINE.json configs[2]: N=16384 observations (32768 x 32768 fp64 covariance).  The oracle
    cannot factorise this in test time, so parity rests on size-independent properties: the GP
    identity K alpha = y - noise alpha at observation sites (ties covariance build, Cholesky, L^-1,
    alpha and the predictive mean together), the interpolation property of the variance, bounds,
    grid-partition invariance, and agreement of the two likelihood entry points."""
    N = 16384
    X, y = synthetic.drifter_snapshot(N, config_id=3)
    theta, noise = (1.3, 3.1, 0.2), 0.05
    m = gp.HelmholtzGP(X, y, *theta, noise)
    lml = m.fit()
    assert np.isfinite(lml)
    sel = np.random.default_rng(0).choice(N, 512, replace=False)
    pm, pv = m.predict(X[sel])
    al = m.alpha()
    idx = torch.as_tensor(sel, device=DEV)
    rhs = torch.cat([m.y[idx], m.y[N + idx]]) - noise * torch.cat([al[idx], al[N + idx]])
    torch.testing.assert_close(pm, rhs, rtol=1e-9, atol=1e-11)
    # at an observation site var = k** - k*' (K + noise I)^-1 k*  lies in (0, noise): the data pin the field
    # down to better than the noise level, never exactly
    kss = orc.helmholtz_Kdiag(1, *theta)[0]
    assert float(pv.min()) > 0.0 and float(pv.max()) < noise
    # far outside the observations the prior variance is recovered
    far = X.max(axis=0) + np.array([[60.0, 60.0], [90.0, 10.0]])
    fm, fv = m.predict(far)
    torch.testing.assert_close(fv, torch.full_like(fv, kss), rtol=1e-9, atol=0)
    assert float(fm.abs().max()) < 1e-9
    # partition invariance, bit for bit (one 600-point shard vs two)
    G = synthetic.prediction_grid(X, 30, 20)
    ma, va = m.predict(G)
    m1, v1 = m.predict(G[:217])
    m2, v2 = m.predict(G[217:])
    assert torch.equal(torch.cat([m1[:217], m2[:383], m1[217:], m2[383:]]), ma)
    assert torch.equal(torch.cat([v1[:217], v2[:383], v1[217:], v2[383:]]), va)
    lml2, grad = m.lml_and_grad()
    assert lml2 == lml and np.all(np.isfinite(grad))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process():
    """Function attributes, side streams and SM counts are set up per device: a model on cuda:1 in
    a process that already used cuda:0 gives the same bits."""
    X, y = synthetic.drifter_snapshot(600, config_id=13)
    Xs = synthetic.prediction_grid(X, 40, 30)
    a = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05, device="cuda:0")
    la, ga = a.lml_and_grad()
    ma, va = a.predict(Xs)
    b = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05, device="cuda:1")
    lb, gb = b.lml_and_grad()
    mb, vb = b.predict(Xs)
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert la == lb and np.array_equal(ga, gb)
    assert torch.equal(ma.cpu(), mb.cpu()) and torch.equal(va.cpu(), vb.cpu())
    g = gp.ScalarGP(np.c_[X, X[:, :1]], y[:600], [0.5], [[2.0, 2.0, 3.0]], 0.01, device="cuda:1")
    g.fit()
    assert g.predict(np.c_[Xs, Xs[:, :1]])[0].device.index == 1


def test_results_do_not_depend_on_previous_workspace_contents():
    """Every kernel family: the workspace is scratch, whatever it held before (zeros, NaN patterns,
    huge numbers) the likelihood, its gradient and the prediction come out bit-identical -- i.e. no
    kernel reads a tile it did not write (the upper tiles of the factor are never initialised)."""
    rng = np.random.default_rng(0)
    N = 300
    X = rng.uniform(0, 10, (N, 3))
    y = rng.normal(size=2 * N)
    Xs = rng.uniform(0, 10, (500, 3))
    makers = {
        "helm": (lambda: gp.HelmholtzGP(X[:, 1:], y, 1.3, 3.1, 0.2, 0.05), Xs[:, 1:]),
        "st": (lambda: gp.SpaceTimeGP(X, y, 1.3, 3.1, 0.2, 1.5, 0.8, 0.05), Xs),
        "hsum": (lambda: gp.HelmholtzSumGP(X, y, [0, 1], [[0.6, 1.5, 1.3, 2.0], [1.4, 0.7, 3.1, 2.2]], 0.05), Xs),
        "rbf": (lambda: gp.ScalarGP(X, y[:N], [1.0, 0.5], [[1, 2, 3], [3, 2, 1]], 0.05), Xs),
    }
    for name, (make, P) in makers.items():
        outs = []
        for fill in (0, 0xFF, None):
            g = make()
            if fill is None:
                g.ws.view(torch.float64)[:] = torch.randn(g.ws.numel() // 8, dtype=torch.float64, device=g.ws.device) * 1e30
            else:
                g.ws.fill_(fill)
            l, gr = g.lml_and_grad()
            m, v = g.predict(P)
            outs.append((l, gr.copy(), m.clone(), v.clone()))
        for o in outs[1:]:
            assert o[0] == outs[0][0] and np.array_equal(o[1], outs[0][1]), name
            assert torch.equal(o[2], outs[0][2]) and torch.equal(o[3], outs[0][3]), name


def test_refined_prediction_all_families():
    """engine.refined_predict: on a well-conditioned problem it reproduces the fused kernel (and the
    oracle); on an ill-conditioned one (noise 1e-9 under a unit prior variance, cond ~ 1e11) it stays
    with the oracle's backward-stable solve where the fused explicit-inverse path drifts."""
    rng = np.random.default_rng(2)
    N, M = 260, 333
    X3 = np.c_[rng.uniform(0, 4, N), rng.uniform(0, 10, (N, 2))]
    Xs3 = np.c_[rng.uniform(0, 4, M), rng.uniform(0, 10, (M, 2))]
    y = np.concatenate([np.sin(X3[:, 1] / 2), np.cos(X3[:, 2] / 2)]) + 0.01 * rng.normal(size=2 * N)
    ty, pr = [0, 1], [[0.6, 1.5, 1.3, 2.0], [1.4, 0.7, 3.1, 2.2]]
    # ill-conditioned case: the data carry 1e-2 of noise against a modelled 1e-9, so |alpha| ~ 1e7 and even a
    # backward-stable solve is only good to eps n |K| |alpha| ~ 5e-7: that floor is the tolerance there
    for noise, tol_fused, tol_mean, tol_var in ((0.05, 1e-9, 2e-8, (1e-6, 1e-10)), (1e-9, None, 5e-6, (1e-3, 1e-9))):
        cases = [
            (gp.HelmholtzGP(X3[:, 1:], y, 1.3, 3.1, 0.2, noise), Xs3[:, 1:],
             lambda: orc.predict(X3[:, 1:], orc.fit(X3[:, 1:], y, 1.3, 3.1, 0.2, noise), 1.3, 3.1, 0.2, Xs3[:, 1:])),
            (gp.SpaceTimeGP(X3, y, 1.3, 3.1, 0.2, 1.5, 0.8, noise), Xs3,
             lambda: orc.st_predict(X3, orc.st_fit(X3, y, 1.3, 3.1, 0.2, 1.5, 0.8, noise), 1.3, 3.1, 0.2, 1.5, 0.8, Xs3)),
            (gp.HelmholtzSumGP(X3, y, ty, pr, noise), Xs3,
             lambda: orc.hsum_predict(X3, orc.hsum_fit(X3, y, ty, pr, noise), ty, pr, Xs3)),
            (gp.ScalarGP(X3, y[:N], [1.0, 0.5], [[1, 2, 3], [3, 2, 1]], noise), Xs3,
             lambda: orc.rbf_predict(X3, orc.rbf_fit(X3, y[:N], [1.0, 0.5], [[1, 2, 3], [3, 2, 1]], noise), [1.0, 0.5], [[1, 2, 3], [3, 2, 1]], Xs3)),
        ]
        lmls = [lambda: orc.fit(X3[:, 1:], y, 1.3, 3.1, 0.2, noise)["lml"],
                lambda: orc.st_fit(X3, y, 1.3, 3.1, 0.2, 1.5, 0.8, noise)["lml"],
                lambda: orc.hsum_fit(X3, y, ty, pr, noise)["lml"],
                lambda: orc.rbf_fit(X3, y[:N], [1.0, 0.5], [[1, 2, 3], [3, 2, 1]], noise)["lml"]]
        for (g, P, oracle), lml_oracle in zip(cases, lmls):
            mo, vo = oracle()
            mr, vr = g.predict_refined(P, chunk_elems=60000)          # several chunks
            scale = max(np.abs(mo).max(), 1e-3)
            np.testing.assert_allclose(mr.cpu().numpy(), mo, rtol=0, atol=tol_mean * scale)
            np.testing.assert_allclose(vr.cpu().numpy(), vo, rtol=tol_var[0], atol=tol_var[1])
            assert (g.cond_bound() > 1e7) == (noise < 1e-6)
            # the likelihood keeps its 1e-6 bar in both regimes (robust factorisation in the second)
            lo = lml_oracle()
            assert abs(g.fit() - lo) <= 1e-6 * abs(lo), (type(g).__name__, noise)
            if tol_fused is not None:
                mf, vf = g.predict(P)
                np.testing.assert_allclose(mf.cpu().numpy(), mr.cpu().numpy(), rtol=0, atol=tol_fused * scale)
                np.testing.assert_allclose(vf.cpu().numpy(), vr.cpu().numpy(), rtol=1e-8, atol=1e-11)
