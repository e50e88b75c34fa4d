"""Synthetic drifter tracks for the benchmark / parity configurations.

Host-side input generation only (the reference's LASER pickles are not shipped).  A
jittered lattice at the seeding density of simulTracks.pkl (0.5 km spacing), velocities from
a Helmholtz mix of a stream-function bump and a potential bump in the spirit of
GP_scripts.generate_2D_gaussian (GP_scripts.py:202-223), plus observation noise.
"""
from __future__ import annotations

import math

import numpy as np

BASE_SEED = 20160207


def drifter_snapshot(N, config_id=2, seed_offset=0, spacing=0.5, jitter=0.2, noise_sd=0.05):
    """X[N,2] km, y[2N] = [u; v] m/s."""
    rng = np.random.default_rng(BASE_SEED + config_id + 1000003 * seed_offset)
    side = int(math.ceil(math.sqrt(N)))
    gx, gy = np.meshgrid(np.arange(side) * spacing, np.arange(side) * spacing)
    P = np.stack([gx.reshape(-1), gy.reshape(-1)], axis=1)[:N]
    P = P + rng.uniform(-jitter, jitter, size=P.shape)
    L = side * spacing
    cx, cy = 0.45 * L, 0.55 * L
    w = max(L / 4.0, 2.0)
    dx, dy = (P[:, 0] - cx) / w, (P[:, 1] - cy) / w
    g = np.exp(-(dx * dx + dy * dy))
    amp = 0.4 * w / math.sqrt(2.0) * math.exp(0.5)       # peak |u| ~ 0.4 m/s
    # stream function psi = amp*g -> (u,v) = (dpsi/dy, -dpsi/dx); potential phi = 0.5*amp*g -> grad phi
    u = amp * (-2 * dy / w) * g + 0.5 * amp * (-2 * dx / w) * g
    v = -amp * (-2 * dx / w) * g + 0.5 * amp * (-2 * dy / w) * g
    u = u + rng.normal(0.0, noise_sd, size=N)
    v = v + rng.normal(0.0, noise_sd, size=N)
    return P, np.concatenate([u, v])


def prediction_grid(X, nx, ny, margin=2.0):
    """Regular nx x ny grid covering the observations + margin; rows ordered (y, x) like the
    reference's meshgrid/reshape (GP_laser.py:147-152; krig.py:670-678)."""
    x = np.linspace(X[:, 0].min() - margin, X[:, 0].max() + margin, nx)
    y = np.linspace(X[:, 1].min() - margin, X[:, 1].max() + margin, ny)
    Xg, Yg = np.meshgrid(x, y)
    return np.stack([Xg.reshape(-1), Yg.reshape(-1)], axis=1)
