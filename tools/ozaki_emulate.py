"""CPU emulation of the int8-sliced (Ozaki) predictive product  V = Z K*  (Z = L^-1), to decide the
slice count before writing the tcgen05 kernel.  Integers are exact in float64 here (sums < 2^53), so this
reproduces bit for bit what int32 accumulators + an fp64 recombination produce on the device.

    python tools/ozaki_emulate.py [N] [M] [noise]
"""
import sys
import os
import time

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gp_oracle as orc            # noqa: E402
import importlib.util                          # noqa: E402

spec = importlib.util.spec_from_file_location("syn", os.path.join(os.path.dirname(__file__), "..", "2d-gp_b200", "synthetic.py"))
syn = importlib.util.module_from_spec(spec)
spec.loader.exec_module(syn)


def slice_rows(A, s, axis_scale):
    """Signed base-256 digits of A scaled per row (axis_scale = 1: one power of two per row; 0: per column;
    None: one global).  Returns digits [s, ...] (float64 holding integers in [-128, 127]) and the scale (value
    of one unit of the LAST digit)."""
    if axis_scale is None:
        mx = np.abs(A).max()
        mx = np.full((1, 1), mx)
    else:
        mx = np.abs(A).max(axis=axis_scale, keepdims=True)
    mx = np.where(mx == 0, 1.0, mx)
    e = np.ceil(np.log2(mx)) + 1                 # |A| / 2^e <= 0.5  -> top digit within [-64, 64]
    nbits = 8 * s - 1
    q = np.rint(A * np.exp2(nbits - e)).astype(np.int64)      # integer of up to 8s-1 bits (+sign), s <= 7
    unit = np.exp2(e - nbits)
    digs = []
    for p in range(s):                           # least significant digit first
        d = ((q + 128) & 255) - 128
        digs.append(d.astype(np.float64))
        q = (q - d) >> 8
    assert np.all(q == 0)
    return np.array(digs[::-1]), unit            # most significant first


def sliced_product(Zd, zu, Kd, ku, s, keep):
    """sum over digit pairs (i, j) with i + j < keep of Z_i K_j 256^(2s-2-i-j), times the units."""
    n, m = Zd.shape[1], Kd.shape[2]
    acc = np.zeros((n, m))
    for d in range(keep - 1, -1, -1):            # small terms first
        P = np.zeros((n, m))
        for i in range(d + 1):
            j = d - i
            if i < s and j < s:
                P += Zd[i] @ Kd[j]
        acc += P * 256.0 ** (2 * s - 2 - d)
    return acc * zu * ku


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    noise = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
    theta = (1.3, 3.1, 0.2)
    X, y = syn.drifter_snapshot(N, config_id=2)
    Xs_all = syn.prediction_grid(X, 320, 320)
    Xs = Xs_all[np.random.default_rng(0).choice(Xs_all.shape[0], M, replace=False)]
    K = orc.helmholtz_K(X, None, *theta) + noise * np.eye(2 * N)
    L = sla.cholesky(K, lower=True)
    Z = sla.solve_triangular(L, np.eye(2 * N), lower=True)
    Ks = orc.helmholtz_K(X, Xs, *theta)          # [2N, 2M]
    kss = orc.helmholtz_Kdiag(M, *theta)
    V = Z @ Ks
    var_ref = kss - (V * V).sum(0)
    # fp64 reference in extended precision for the truth
    Vl = Z.astype(np.longdouble) @ Ks.astype(np.longdouble)
    var_true = (kss.astype(np.longdouble) - (Vl * Vl).sum(0)).astype(np.float64)
    print("N=%d n=%d M=%d noise=%g  var in [%.3e, %.3e]  fp64-vs-longdouble rel %.2e" %
          (N, 2 * N, M, noise, var_true.min(), var_true.max(), np.abs(var_ref / var_true - 1).max()))
    for s in (5, 6, 7):
        for kscale in ("global",):
            t0 = time.time()
            Zd, zu = slice_rows(Z, s, 1)
            Kd, ku = slice_rows(Ks, s, 0 if kscale == "col" else None)
            for keep in (s, s + 1):
                Vs = sliced_product(Zd, zu, Kd, ku, s, keep)
                var = kss - (Vs * Vs).sum(0)
                print("  s=%d K*-scale=%-6s pairs i+j<%d (%2d products): max|dV|/max|V| %.2e   var rel err max %.2e   (%.1fs)" %
                      (s, kscale, keep, sum(min(d, s - 1) - max(0, d - s + 1) + 1 for d in range(keep)),
                       np.abs(Vs - V).max() / np.abs(V).max(), np.abs(var / var_true - 1).max(), time.time() - t0))


if __name__ == "__main__":
    main()
