"""Job-array driver with the reference's argv contract (runPredict.py:1-48):

    python -m gp2d_b200.runPredict <1-based job index>

The reference maps the index to a prediction time (idt0 = 8, 12, 16, 20 in units of
tg = arange(12, 36, 0.5)) and a velocity component, then calls krig.scikit_prior (scalar
RBF in scikit-learn).  The vector-valued kernel predicts both components at once, so the
component half of the index range repeats the same work; the grid limits are the reference's.
"""
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
    print('Part ', ind + 1, ' of ', idt0.size, '.')
    idt = idt0[ind]
    ylim, xlim, dx = [1, 15], [-5, 15], 0.1
    outFile = 'outputs_proposal/skip2_day1/' + 'rbfModel_T' + str(T) + '_dt' + str(dt) + '_nK' + str(nK)
    return krig.predict(outFile, tlim=[tg[idt], tg[idt] + 1], ylim=ylim, xlim=xlim, dt=1, dx=dx)


if __name__ == "__main__":
    main()
