"""GPU tests of the batched entry points (survey kernel k6): gp2d_fit_batched / gp2d_lml_grad_batched must
give every problem of a batch exactly what its single-problem call gives -- bit for bit -- for shared data
(restarts, krig.py:450; GP_plots.py:765) and per-problem data (time slices, krig.py:541-557).
Run on the B200 box:  pytest tests -m gpu"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():          # collected on the CPU box, run on the GPU box
    pytest.skip("no CUDA device", allow_module_level=True)

import gp2d_b200 as gp                      # noqa: E402
from gp2d_b200 import models, myKernel, synthetic   # noqa: E402
from oracle import gp_oracle as orc        # noqa: E402

THETA4 = np.array([[2.0, 2.0, 0.5, 0.05], [1.3, 3.1, 0.2, 0.05], [0.6, 0.6, 1.0, 0.1], [0.7, 1.9, 0.0, 0.02],
                   [1.0, 1.0, 0.5, 0.3], [3.0, 0.9, 0.8, 0.01], [0.45, 2.2, 0.35, 0.07]])


def single(X, y, th, jitter=0.0, compat=False, grad=True):
    m = gp.HelmholtzGP(X, y, th[0], th[1], th[2], th[3], jitter=jitter)
    if grad:
        lml, g = m.lml_and_grad(reference_compat=compat)
    else:
        lml, g = m.fit(), None
    return m, lml, g


@pytest.mark.parametrize("N", [130, 300, 1100])      # npad 384 / 640 / 2304: leaf-only levels, side streams, both tile shapes
def test_lml_grad_batched_is_bit_identical_shared_data(N):
    X, y = synthetic.drifter_snapshot(N, config_id=4, seed_offset=N)
    Xs = synthetic.prediction_grid(X, 23, 7)
    B = len(THETA4)
    hb = gp.HelmholtzBatch(X, y, B=B, jitter=1e-8)
    for compat in (False, True):
        lml, grad, info = hb.lml_and_grad(THETA4, reference_compat=compat)
        assert not info.any()
        for b in range(B):
            m, l1, g1 = single(X, y, THETA4[b], jitter=1e-8, compat=compat)
            assert lml[b] == l1, (b, lml[b], l1)
            np.testing.assert_array_equal(grad[b], g1)
            # the workspace of problem b is a complete fit state
            mb, vb = hb.predict(b, Xs)
            m1, v1 = m.predict(Xs)
            assert torch.equal(mb, m1) and torch.equal(vb, v1)
    # a partial round (fewer live restarts than workspaces) uses the first nb workspaces
    lml3, grad3, _ = hb.lml_and_grad(THETA4[2:5], nb=3)
    np.testing.assert_array_equal(lml3, hb.lml_and_grad(THETA4)[0][2:5])


def test_fit_batched_is_bit_identical_per_problem_data():
    N, B = 260, 5
    Xb = np.stack([synthetic.drifter_snapshot(N, config_id=5, seed_offset=s)[0] for s in range(B)])
    yb = np.stack([synthetic.drifter_snapshot(N, config_id=5, seed_offset=s)[1] for s in range(B)])
    hb = gp.HelmholtzBatch(Xb, yb, jitter=0.0)
    alpha = torch.empty((B, 2 * N), dtype=torch.float64, device="cuda")
    hb.fit_async(THETA4[:B], alpha_out=alpha)
    lml = hb._out.reshape(-1)[:B].cpu().numpy()
    for b in range(B):
        m = gp.HelmholtzGP(Xb[b], yb[b], *THETA4[b], jitter=0.0)
        a1 = torch.empty(2 * N, dtype=torch.float64, device="cuda")
        m.fit_async(alpha_out=a1)
        assert float(m._scal[0].item()) == lml[b]
        assert torch.equal(alpha[b], a1)
        f = orc.fit(Xb[b], yb[b], *THETA4[b])
        assert abs(lml[b] - f["lml"]) <= 1e-6 * abs(f["lml"])
        np.testing.assert_allclose(alpha[b].cpu().numpy(), f["alpha"], rtol=1e-8, atol=1e-9 * np.abs(f["alpha"]).max())


def test_batch_with_ill_conditioned_and_indefinite_members():
    """A robust-mode member (tiny noise: refined factorisation, runs one by one) and a member that is
    not positive definite sit between plain members; everyone still gets its single-problem result."""
    N = 200
    X, y = synthetic.drifter_snapshot(N, config_id=4)
    X = X.copy()
    th = THETA4[:5].copy()
    th[1, 3] = 1e-9                  # cond bound 2N k** / noise ~ 1e11 > 1e7: robust mode
    hb = gp.HelmholtzBatch(X, y, B=5, jitter=0.0)
    lml, grad, info = hb.lml_and_grad(th)
    for b in range(5):
        _, l1, g1 = single(X, y, th[b])
        assert lml[b] == l1
        np.testing.assert_array_equal(grad[b], g1)
    # duplicate observation points and zero noise: singular covariance
    Xd = X.copy()
    Xd[1] = Xd[0]
    th[1, 3] = 0.05
    th[3, 3] = 0.0
    hb2 = gp.HelmholtzBatch(Xd, y, B=5, jitter=0.0)
    lml, info = hb2.fit(th)
    assert info[3] > 0 and not info[[0, 1, 2, 4]].any()
    m = gp.HelmholtzGP(Xd, y, *th[0])
    assert m.fit() == lml[0]


def test_krig_snapshots_matches_oracle_and_single_calls():
    N, S = 220, 5
    Xb = np.stack([synthetic.drifter_snapshot(N, config_id=5, seed_offset=s)[0] for s in range(S)])
    yb = np.stack([synthetic.drifter_snapshot(N, config_id=5, seed_offset=s)[1] for s in range(S)])
    Xs = synthetic.prediction_grid(Xb[0], 17, 9)
    th = (1.3, 3.1, 0.2, 0.05)
    mean, var, lml = gp.krig_snapshots(Xb, yb, Xs, *th, batch=2)         # 2 + 2 + 1: a ragged last batch
    for s in range(S):
        f = orc.fit(Xb[s], yb[s], *th)
        mo, vo = orc.predict(Xb[s], f, *th[:3], Xs)
        np.testing.assert_allclose(mean[s], mo, rtol=1e-8, atol=1e-9 * np.abs(mo).max())
        np.testing.assert_allclose(var[s], vo, rtol=1e-8, atol=1e-12)
        assert abs(lml[s] - f["lml"]) <= 1e-6 * abs(f["lml"])
        m1, v1, l1 = gp.fit_predict_host(Xb[s], yb[s], *th, Xs)
        np.testing.assert_array_equal(mean[s], m1)
        np.testing.assert_array_equal(var[s], v1)
        assert lml[s] == l1


def test_lockstep_restarts_reproduce_the_sequential_runs():
    """optimize_restarts in lock step (one batched evaluation per round) must produce the very runs of the
    one-at-a-time driver: same start points, same objective values bit for bit, hence same iterates."""
    X, y = synthetic.drifter_snapshot(150, config_id=4)

    def model():
        return models.GPRegression(X, y[:, None], myKernel.myKernel(2, [0, 1], 1.0, 1.0, 0.5), noise_var=0.1)
    a, b = model(), model()
    ra = a.optimize_restarts(num_restarts=5, verbose=False, seed=11, max_iters=25, batched=False)
    rb = b.optimize_restarts(num_restarts=5, verbose=False, seed=11, max_iters=25, batched=True)
    assert len(ra) == len(rb) == 5
    for u, v in zip(ra, rb):
        assert u.f_opt == v.f_opt and u.funct_eval == v.funct_eval
        np.testing.assert_array_equal(u.x_opt, v.x_opt)
    np.testing.assert_array_equal(a.param_array, b.param_array)
    assert b.lockstep_stats["evaluations"] == sum(r.funct_eval for r in rb)
    assert b.lockstep_stats["rounds"] < b.lockstep_stats["evaluations"]
    # sharded over two "ranks" in lock step: the union of the runs is the same set
    c0, c1 = model(), model()
    r0 = c0.optimize_restarts(num_restarts=5, verbose=False, seed=11, max_iters=25, rank=0, world=2, batched=True)
    r1 = c1.optimize_restarts(num_restarts=5, verbose=False, seed=11, max_iters=25, rank=1, world=2, batched=True)
    assert sorted(r.f_opt for r in r0 + r1) == sorted(r.f_opt for r in ra)


def test_batched_bad_arguments():
    from gp2d_b200._lib import lib
    assert lib.gp2d_fit_batched_workspace_bytes(100, 3) == 3 * lib.gp2d_fit_workspace_bytes(100)
    assert lib.gp2d_fit_batched_workspace_bytes(100, 0) == 0
    X, y = synthetic.drifter_snapshot(64, config_id=4)
    hb = gp.HelmholtzBatch(X, y, B=2)
    with pytest.raises(gp.engine.Gp2dError):
        hb.fit(np.array([[1.0, 1.0, 0.5, 0.05], [-1.0, 1.0, 0.5, 0.05]]))       # invalid theta -> argument 7
    with pytest.raises(ValueError):
        hb.fit(np.ones((3, 4)))
