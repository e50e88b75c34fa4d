"""How many of the S (S + 1) / 2 digit-slice products per k-step are identically zero?  (numpy emulation)

A balanced base-256 digit slice of a tile is all zero when every entry of the tile is small against the scale of
its row (Z) or of the panel (K*): entries far from the diagonal of Z = L^-1 and covariances between distant points.
Counts, for a configuration, the MMAs the int8 predictive kernel would issue if it skipped all-zero slices, for the
observation order as given and for spatially blocked orders.
    python tools/sparsity_emulate.py [N] [grid side] [order: none|morton|block2|random]"""
import os
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
spec = importlib.util.spec_from_file_location("syn", os.path.join(os.path.dirname(__file__), "..", "2d-gp_b200", "synthetic.py"))
syn = importlib.util.module_from_spec(spec); spec.loader.exec_module(syn)
from oracle import gp_oracle as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
side = int(sys.argv[2]) if len(sys.argv) > 2 else 320
order = sys.argv[3] if len(sys.argv) > 3 else "none"
S, SMAX, NG = 6, 7, 40
theta, noise = (1.3, 3.1, 0.2), 0.05
X, y = syn.drifter_snapshot(N, config_id=2)
Xs = syn.prediction_grid(X, side, side)


def morton(P, bits=10):
    q = ((P - P.min(0)) / (np.ptp(P, axis=0).max() + 1e-12) * ((1 << bits) - 1)).astype(np.uint64)
    def spread(v):
        r = np.zeros_like(v)
        for b in range(bits):
            r |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(2 * b)
        return r
    return spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))


if order == "morton":
    perm = np.argsort(morton(X), kind="stable")
elif order.startswith("block"):
    # groups of 16 observations that form compact patches: sort by (patch row, patch column, inside)
    w = float(order[5:] or 2.0)           # patch edge in km
    perm = np.lexsort((X[:, 0], X[:, 1], np.floor(X[:, 0] / w), np.floor(X[:, 1] / w)))
elif order == "random":
    perm = np.random.default_rng(1).permutation(N)
else:
    perm = np.arange(N)
X = X[perm]; y = np.concatenate([y[:N][perm], y[N:][perm]])

n = 2 * N
npad = (n + 127) // 128 * 128
K = orc.helmholtz_K(X, X, *theta)
# pair-interleave
idx = np.empty(n, dtype=np.int64); idx[0::2] = np.arange(N); idx[1::2] = N + np.arange(N)
Ki = K[np.ix_(idx, idx)] + noise * np.eye(n)
L = np.linalg.cholesky(Ki)
Z = sla.solve_triangular(L, np.eye(n), lower=True)
Zp = np.eye(npad); Zp[:n, :n] = Z


def digits_needed(q):
    """smallest p with -2^(8p-1) <= q < 2^(8p-1) (0 for q == 0)"""
    a = np.abs(q)
    p = np.zeros(a.shape, dtype=np.int64)
    nz = a >= 0.5
    p[nz] = np.floor((np.log2(a[nz]) + 1.0) / 8.0).astype(np.int64) + 1
    return p


# Z: per-row scale from the row max over its tile columns; q = rint(z qs), |q| <= 2^(8 SMAX - 2)
rowmax = np.abs(Zp).max(axis=1)
ex = np.frexp(rowmax)[1] + 1
qs = np.ldexp(1.0, (8 * SMAX - 1) - ex)
Q = np.rint(Zp * qs[:, None])
pz = digits_needed(Q)                         # digits needed per entry, out of SMAX
nb = npad // 128
imin = np.full((nb, npad // 32), S, dtype=np.int64)      # leading all-zero slices of tile (rb, ks)
for rb in range(nb):
    for ks in range(4 * (rb + 1)):
        t = pz[rb * 128:(rb + 1) * 128, ks * 32:(ks + 1) * 32]
        imin[rb, ks] = min(S, SMAX - t.max())
print("Z tiles: leading zero slices histogram (lower triangle only):",
      np.bincount(np.concatenate([imin[rb, :4 * (rb + 1)] for rb in range(nb)]), minlength=S + 1))

# K*: global scale from k** = kss
kss = orc.helmholtz_Kdiag(1, *theta)[0]
exk = np.frexp(kss)[1] + 1
kscale = np.ldexp(1.0, (8 * S - 1) - exk)
M = Xs.shape[0]
ntile = (M + NG - 1) // NG
rng = np.random.default_rng(0)
tiles = np.sort(rng.choice(ntile, size=min(ntile, 256), replace=False))
nks = npad // 32
tot_dense = 0; tot_b = 0; tot_ab = 0; ksteps_all = 0; ksteps_live = 0
pairs = [(i, j) for i in range(S) for j in range(S) if i + j < S]
jhist = np.zeros(S + 1, dtype=np.int64)
for t in tiles:
    g = Xs[t * NG:(t + 1) * NG]
    Ks = orc.helmholtz_K(X, g, *theta)                      # [2N, 2G] component-major
    G = g.shape[0]
    A = np.abs(Ks).reshape(2, N, 2, G).max(axis=(0, 2, 3))  # per observation
    Ap = np.zeros(npad // 2); Ap[:N] = A
    kmax = Ap.reshape(nks, 16).max(axis=1)                   # per k-step (16 observations)
    pk = digits_needed(np.rint(kmax * kscale))
    jmin = np.minimum(S, S - pk)
    jhist += np.bincount(jmin, minlength=S + 1)
    for rb in range(nb):
        k1 = 4 * (rb + 1)
        im = imin[rb, :k1]; jm = jmin[:k1]
        tot_dense += len(pairs) * k1
        nb_only = np.zeros(k1, dtype=np.int64); nab = np.zeros(k1, dtype=np.int64)
        for (i, j) in pairs:
            nb_only += (j >= jm)
            nab += (j >= jm) & (i >= im)
        tot_b += nb_only.sum(); tot_ab += nab.sum()
        ksteps_all += k1; ksteps_live += (nab > 0).sum()
print("K* k-step tiles: leading zero slices histogram:", jhist)
print("N=%d grid %dx%d order=%s: MMAs issued / dense: K* zeros only %.3f, K* and Z zeros %.3f; k-steps with any work %.3f" %
      (N, side, side, order, tot_b / tot_dense, tot_ab / tot_dense, ksteps_live / ksteps_all))
