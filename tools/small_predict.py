"""Predict time of small problems: fp64 kernel (mode 1) against the int8 kernel (mode 6 / automatic)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp2d_b200 as gp
from gp2d_b200 import synthetic as syn
for (N, side) in ((100, 20), (400, 51), (400, 160), (1000, 51), (1000, 200), (2000, 100)):
    X, y = syn.drifter_snapshot(N, config_id=1)
    Xsd = gp.as_dev(syn.prediction_grid(X, side, side))
    out = []
    for mode in (1, 6):
        gp.set_predict_i8(mode)
        m = gp.HelmholtzGP(X, y, 2.0, 2.0, 0.5, 0.05)
        m.fit()
        for _ in range(5):
            m.predict(Xsd)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            m.predict(Xsd)
        b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / 20)
    print("N=%5d grid %3dx%-3d: fp64 kernel %.3f ms, int8 kernel %.3f ms" % (N, side, side, out[0], out[1]), flush=True)
gp.set_predict_i8(0)
