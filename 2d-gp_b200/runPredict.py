"""Job-array driver with the reference's argv contract (runPredict.py:1-48):

    python -m gp2d_b200.runPredict <1-based job index>

The index selects a prediction time (idt0 = 8, 12, 16, 20 in units of tg = arange(12, 36, 0.5))
and a velocity component (first half of the index range 'v', second half 'u'), then calls
krig.scikit_prior with the reference's grid limits -- the scalar ARD-RBF model rebuilt with the
optimised hyper-parameters and predicted on the GPU.  GP2D_PREDICT_DIR overrides the directory
of the model files (the reference hard-codes 'outputs_proposal/skip2_day1/')."""
import os
import sys

import numpy as np


def main(argv=None):
    from . import krig
    argv = sys.argv if argv is None else argv
    ind = int(argv[1]) - 1
    T, dt, nK = 1.0, 2, 2
    idt0 = np.arange(8, 24, 4)
    tg = np.arange(12, 36, 0.5)
    if ind >= idt0.size:
        ind = ind - idt0.size
        var = 'u'
    else:
        var = 'v'
    print('Part ', ind + 1, ' of ', idt0.size, '.')
    idt = idt0[ind]
    ylim, xlim, dx = [1, 15], [-5, 15], 0.1
    outDir = os.environ.get("GP2D_PREDICT_DIR", 'outputs_proposal/skip2_day1/')
    outFile = outDir + '/' + 'rbfModel_T' + str(T) + '_dt' + str(dt) + '_nK' + str(nK)
    return krig.scikit_prior(outFile, varname=var, dt=tg[idt], tlim=8, radar='', xlim=xlim, ylim=ylim, dx=dx)


if __name__ == "__main__":
    main()
