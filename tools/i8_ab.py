"""A/B timing of two bring-up settings of the int8 predictive kernel in thermal steady state: the settings alternate,
each block runs ~1.5 s, times in ms (what counts) and in SM clocks (what the kernel does).
    python tools/i8_ab.py dbgA dbgB [N] [grid side] [rounds]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp2d_b200 as gp                                   # noqa: E402
from gp2d_b200 import synthetic as syn                   # noqa: E402
from gp2d_b200._lib import lib                           # noqa: E402

lib.gp2d_dbg_set_i8.restype = C.c_int
lib.gp2d_dbg_set_i8.argtypes = [C.c_int]
lib.gp2d_dbg_i8_counters6.restype = C.c_int
lib.gp2d_dbg_i8_counters6.argtypes = [C.POINTER(C.c_ulonglong)]
cnt = (C.c_ulonglong * 8)()
A, B = int(sys.argv[1]), int(sys.argv[2])
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
side = int(sys.argv[4]) if len(sys.argv) > 4 else 320
rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 4
X, y = syn.drifter_snapshot(N, config_id=2)
Xsd = gp.as_dev(syn.prediction_grid(X, side, side))
gp.set_predict_i8(6)
m = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05)
m.fit()
for _ in range(40):
    m.predict(Xsd)
torch.cuda.synchronize()
tot = {A: [], B: []}
for r in range(rounds):
    for dbg in (A, B):
        lib.gp2d_dbg_set_i8(dbg)
        m.predict(Xsd)
        lib.gp2d_dbg_i8_counters6(cnt)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(60):
            m.predict(Xsd)
        b.record()
        torch.cuda.synchronize()
        lib.gp2d_dbg_i8_counters6(cnt)
        ms = a.elapsed_time(b) / 60
        mclk = cnt[4] / max(1, cnt[5]) / 1e6
        tot[dbg].append(ms)
        print("round %d dbg %5d: %7.3f ms  %6.2f Mclk/CTA  -> %.0f MHz" % (r, dbg, ms, mclk, mclk / ms * 1e3), flush=True)
print("mean ms: dbg %d %.3f, dbg %d %.3f" % (A, sum(tot[A]) / rounds, B, sum(tot[B]) / rounds))
lib.gp2d_dbg_set_i8(0)
