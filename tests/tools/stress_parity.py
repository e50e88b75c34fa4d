"""Randomised parity sweep (not part of the test suite): many random sizes / hyper-parameters for the
four kernel families, CUDA path against the CPU oracle at the BASELINE.json tolerances.
Usage: python tests/tools/stress_parity.py [cases] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gp2d_b200 as gp
from oracle import gp_oracle as orc

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = {"mean": 0.0, "var": 0.0, "lml": 0.0, "grad": 0.0}
worst_ill = dict(worst)          # cases in the robust regime (conditioning bound > 1e7)
fails = 0


def rel(a, b, floor):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


for c in range(cases):
    fam = c % 4
    N = int(rng.choice([1, 2, 3, 15, 16, 17, 31, 33, 63, 64, 65, 100, 127, 128, 129, 200, 255, 257, 400, 511, 513, 700]))
    M = int(rng.choice([1, 2, 63, 64, 65, 127, 128, 129, 300, 1000, 2601, 5000]))
    noise = float(10 ** rng.uniform(-3, -1)) if rng.uniform() < 0.8 else float(10 ** rng.uniform(-8, -4))   # robust mode
    side = max(2.0, 0.5 * np.sqrt(N))
    try:
        if fam == 0:
            X = rng.uniform(0, side, (N, 2)); Xs = rng.uniform(-1, side + 1, (M, 2))
            y = rng.normal(size=2 * N) * 0.3
            th = (float(rng.uniform(0.4, 3)), float(rng.uniform(0.4, 3)), float(rng.choice([0.0, 1.0, rng.uniform(0.05, 0.95)])))
            if rng.uniform() < 0.3:
                th = (th[0], th[0], th[2])
            g = gp.HelmholtzGP(X, y, *th, noise, jitter=1e-8)
            lml, grad = g.lml_and_grad()
            mean, var = g.predict(Xs)
            lo, go = orc.lml_and_grad(X, y, *th, noise, jitter=1e-8)
            f = orc.fit(X, y, *th, noise, jitter=1e-8)
            mo, vo = orc.predict(X, f, *th, Xs)
        elif fam == 1:
            D, Q = int(rng.integers(1, 5)), int(rng.integers(1, 5))
            X = rng.uniform(0, side, (N, D)); Xs = rng.uniform(-1, side + 1, (M, D))
            y = rng.normal(size=N) * 0.3
            v, ls = rng.uniform(0.1, 2, Q), rng.uniform(0.5, 4, (Q, D))
            g = gp.ScalarGP(X, y, v, ls, noise, jitter=1e-8)
            lml, grad = g.lml_and_grad()
            mean, var = g.predict(Xs)
            lo, go = orc.rbf_lml_and_grad(X, y, v, ls, noise, jitter=1e-8)
            f = orc.rbf_fit(X, y, v, ls, noise, jitter=1e-8)
            mo, vo = orc.rbf_predict(X, f, v, ls, Xs)
        elif fam == 3:
            ldx, Q = int(rng.choice([2, 3])), int(rng.integers(1, 9))
            X = np.c_[rng.uniform(0, 6, N), rng.uniform(0, side, (N, 2))][:, 3 - ldx:]
            Xs = np.c_[rng.uniform(0, 6, M), rng.uniform(-1, side + 1, (M, 2))][:, 3 - ldx:]
            X, Xs = np.ascontiguousarray(X), np.ascontiguousarray(Xs)
            y = rng.normal(size=2 * N) * 0.3
            ty = rng.integers(0, 2, Q)
            pr = np.c_[rng.uniform(0.05, 1.5, Q), rng.uniform(0.5, 5, Q), rng.uniform(0.4, 3, (Q, 2))]
            g = gp.HelmholtzSumGP(X, y, ty, pr, noise, jitter=1e-8)
            lml, grad = g.lml_and_grad()
            mean, var = g.predict(Xs)
            lo, go = orc.hsum_lml_and_grad(X, y, ty, pr, noise, jitter=1e-8)
            f = orc.hsum_fit(X, y, ty, pr, noise, jitter=1e-8)
            mo, vo = orc.hsum_predict(X, f, ty, pr, Xs)
        else:
            X = np.c_[rng.uniform(0, 6, N), rng.uniform(0, side, (N, 2))]
            Xs = np.c_[rng.uniform(0, 6, M), rng.uniform(-1, side + 1, (M, 2))]
            y = rng.normal(size=2 * N) * 0.3
            th = (float(rng.uniform(0.4, 3)), float(rng.uniform(0.4, 3)), float(rng.uniform(0, 1)),
                  float(rng.uniform(0.3, 3)), float(rng.uniform(0.5, 5)))
            g = gp.SpaceTimeGP(X, y, *th, noise, jitter=1e-8)
            lml, grad = g.lml_and_grad()
            mean, var = g.predict(Xs)
            lo, go = orc.st_lml_and_grad(X, y, *th, noise, jitter=1e-8)
            f = orc.st_fit(X, y, *th, noise, jitter=1e-8)
            mo, vo = orc.st_predict(X, f, *th, Xs)
    except np.linalg.LinAlgError as e:
        print("case %d fam %d N=%d: not positive definite on one side (%s)" % (c, fam, N, e))
        continue
    e_mean = rel(mean.cpu().numpy(), mo, max(1e-3, 1e-1 * np.abs(mo).max()))
    e_var = rel(var.cpu().numpy(), vo, 1e-4)
    e_lml = abs(lml - lo) / max(abs(lo), 1.0)
    e_grad = rel(grad, go, max(1e-1, 1e-1 * np.abs(go).max()))
    if g.cond_bound() <= 1e7:
        for k, v in (("mean", e_mean), ("var", e_var), ("lml", e_lml), ("grad", e_grad)):
            worst[k] = max(worst[k], v)
    else:
        for k, v in (("mean", e_mean), ("var", e_var), ("lml", e_lml), ("grad", e_grad)):
            worst_ill[k] = max(worst_ill[k], v)
    # ill-conditioned draws (noise far below the scatter of the data): |alpha| is huge and even a
    # backward-stable solve is only good to ~ eps n k** |alpha|; that floor, not 1e-8, is the bar there
    n_sc = f["alpha"].size
    kss = g.cond_bound() * (noise + 1e-8) / n_sc
    floor = 100 * 2.2e-16 * n_sc * kss * float(np.abs(f["alpha"]).max()) / max(1e-3, 1e-1 * float(np.abs(mo).max()))
    slack = max(1.0, floor / 1e-8, g.cond_bound() / 1e9)
    if e_mean > 1e-8 * slack or e_var > 1e-8 * slack or e_lml > 1e-6 or e_grad > 1e-6 * slack:
        fails += 1
        print("FAIL case %d fam %d N=%d M=%d noise=%.3g: mean %.2e var %.2e lml %.2e grad %.2e" % (c, fam, N, M, noise, e_mean, e_var, e_lml, e_grad))
print("cases %d, failures %d, worst relative errors %s; robust regime %s" % (
    cases, fails, {k: "%.2e" % v for k, v in worst.items()}, {k: "%.2e" % v for k, v in worst_ill.items()}))
sys.exit(1 if fails else 0)
