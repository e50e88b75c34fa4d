"""Timing of the int8 predictive kernel under its bring-up knobs (wrong results by design for 1, 2, 4).
    python tools/i8_knobs.py [N] [grid side] [slices]"""
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
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
side = int(sys.argv[2]) if len(sys.argv) > 2 else 320
S = int(sys.argv[3]) if len(sys.argv) > 3 else 6
X, y = syn.drifter_snapshot(N, config_id=2)
Xsd = gp.as_dev(syn.prediction_grid(X, side, side))
gp.set_predict_i8(S)
m = gp.HelmholtzGP(X, y, 1.3, 3.1, 0.2, 0.05)
m.fit()
names = {1: "B from one shared panel (L2)", 2: "no epilogue math", 4: "no generation", 8: "dense schedule", 16: "no MMAs", 32: "no copies"}
for dbg in [int(a) for a in (sys.argv[4].split(",") if len(sys.argv) > 4 else "0,8,1,2,4,3,7,9,10,12,15".split(","))]:
    lib.gp2d_dbg_set_i8(dbg)
    m.predict(Xsd)
    torch.cuda.synchronize()
    lib.gp2d_dbg_i8_counters6(cnt)
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); m.predict(Xsd); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    lib.gp2d_dbg_i8_counters6(cnt)
    L = max(1, cnt[3])
    if hasattr(lib, "gp2d_dbg_i8_waitprof"):
        wp = (C.c_ulonglong * 16)()
        lib.gp2d_dbg_i8_waitprof(wp)
        nct = max(1, cnt[5])
        print("      (GP2D_I8_WAITPROF build; line belongs to the next dbg line) wait Mclk per CTA and predict, summed over the warps of a role: epilogue acc_full %.2f | producer panel_full %.2f, empty %.2f | issuers full %.2f, acc_empty %.2f, "
              "first_done %.2f | generators panel_empty %.2f" % tuple(wp[i] / (nct / L * 4) / 1e6 for i in (1, 2, 3, 4, 5, 7, 6)))
    print("dbg %4d: %8.2f ms  %7.2f Mclk/CTA (generators busy %5.2f)  %-40s per launch: %.4g slice products, %.4g stages, %.4g k-steps (%.1f products / stage, %.3f of the k-steps live)" %
          (dbg, min(ts), cnt[4] / max(1, cnt[5]) / 1e6, cnt[6] / max(1, cnt[5]) / 1e6, " + ".join(v for k, v in names.items() if dbg & k) or "production", cnt[0] / L, cnt[1] / L, cnt[2] / L,
           cnt[0] / max(1, cnt[1]), cnt[1] / max(1, cnt[2])), flush=True)
lib.gp2d_dbg_set_i8(0)
gp.set_predict_i8(0)
