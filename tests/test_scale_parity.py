"""Oracle-number parity at the sizes BASELINE.json names (the small-size tests of test_gpu_parity.py use
the same oracle; identities alone covered these sizes before): configs[4] N = 8192 (mean / variance / LML)
and configs[3] N = 2000 (LML gradient, analytic and reference_compat).  The oracle restates
GP_laser.py:113-140 and myKernel.py:27-106; the N = 16 384 case (configs[2]) needs 8.6 GB and a minute of
host Cholesky and lives in tools/config3_parity.py, its result under profiles/.
Run on the B200 box:  pytest tests -m gpu"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():          # collected on the CPU box, run on the GPU box
    pytest.skip("no CUDA device", allow_module_level=True)

import gp2d_b200 as gp                      # noqa: E402
from gp2d_b200 import synthetic             # noqa: E402
from oracle import gp_oracle as orc        # noqa: E402


def test_config5_vs_oracle():
    """One configs[4] snapshot (N = 8192, n = 16 384) against the chunked oracle: LML to 1e-6, mean and
    variance at 300 sampled points of the 100 x 100 grid to 1e-8 (BASELINE.json tolerances)."""
    N = 8192
    theta, noise = (1.3, 3.1, 0.2), 0.05
    X, y = synthetic.drifter_snapshot(N, config_id=5, seed_offset=3)
    grid = synthetic.prediction_grid(X, 100, 100)
    pick = np.random.default_rng(5).choice(grid.shape[0], 300, replace=False)
    m = gp.HelmholtzGP(X, y, *theta, noise)
    lml = m.fit()
    mean, var = m.predict(grid)
    mean, var = mean.cpu().numpy(), var.cpu().numpy()
    M = grid.shape[0]
    f = orc.fit_chunked(X, y, *theta, noise)
    mo, vo = orc.predict(X, f, *theta, grid[pick], chunk=150)
    sel = np.concatenate([pick, M + pick])
    assert abs(lml - f["lml"]) <= 1e-6 * abs(f["lml"]), (lml, f["lml"])
    np.testing.assert_allclose(mean[sel], mo, rtol=1e-8, atol=1e-9 * np.abs(mo).max())
    np.testing.assert_allclose(var[sel], vo, rtol=1e-8, atol=1e-12)
    # the batched entry point at this size: the same numbers bit for bit
    hb = gp.HelmholtzBatch(X[None], y[None], jitter=0.0)
    l2, info = hb.fit([theta + (noise,)])
    assert info[0] == 0 and l2[0] == lml
    m2, v2 = hb.predict(0, grid[pick])
    np.testing.assert_array_equal(m2.cpu().numpy(), mean[sel])
    np.testing.assert_array_equal(v2.cpu().numpy(), var[sel])


@pytest.mark.parametrize("theta", [(2.0, 2.0, 0.5), (1.3, 3.1, 0.2)])
def test_lml_grad_vs_oracle_at_config4_size(theta):
    """configs[3]: log-marginal-likelihood and gradient at N = 2000 (n = 4000), both gradient conventions
    (analytic, and the reference's own integrands myKernel.py:77-81,91-96), GPy's jitter."""
    X, y = synthetic.drifter_snapshot(2000, config_id=4)
    m = gp.HelmholtzGP(X, y, *theta, 0.1, jitter=1e-8)
    for compat in (False, True):
        lml, grad = m.lml_and_grad(reference_compat=compat)
        lo, go = orc.lml_and_grad(X, y, *theta, 0.1, jitter=1e-8, reference_compat=compat)
        assert abs(lml - lo) <= 1e-6 * abs(lo)
        np.testing.assert_allclose(grad, go, rtol=1e-6, atol=1e-6 * np.abs(go).max())
