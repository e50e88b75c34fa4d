"""Job-array driver with the reference's argv contract (runKrig.py:1-36):

    python -m gp2d_b200.runKrig <1-based job index>

The index selects (T, dt, skip, nK) from the same hard-coded tables; the model is built by
krig.kriging with the reference's default kernelType=1 (sum of nK scalar ARD-RBF kernels, one model
per velocity component).  GP2D_KERNEL_TYPE=2|3|4 selects the divergence-free / curl-free /
combined Helmholtz kernel instead; GP2D_LASER=1 reads the LASER pickle instead of the simulations.
"""
import os
import sys

import numpy as np


def main(argv=None):
    from . import krig
    argv = sys.argv if argv is None else argv
    ind = int(argv[1]) - 1
    T = np.array([1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    dt = np.array([1, 1, 2, 2, 2, 1, 1, 2, 2, 2, 1, 1, 2, 2, 2])
    skp = np.array([1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3])
    nK = np.array([1, 2, 1, 2, 3, 1, 2, 1, 2, 3, 1, 2, 1, 2, 3])
    st = 0
    et = st + T[ind] * 24 * 60 // 15
    laser = int(os.environ.get("GP2D_LASER", "0"))       # 1 for the data, 0 for simulations
    outFile = 'rbfModel_T' + str(T[ind]) + '_dt' + str(dt[ind]) + '_nK' + str(nK[ind])
    outDir = ('skip_' if laser == 1 else 'Simulations/skip_') + str(skp[ind])
    os.makedirs(outDir, exist_ok=True)
    outFile = outDir + '/' + outFile
    print(outFile)
    return krig.kriging(st, et, sample_step=-int(dt[ind]), skip=int(skp[ind]), nKernels=int(nK[ind]),
                        output=outFile, laser=laser, kernelType=int(os.environ.get("GP2D_KERNEL_TYPE", "1")))


if __name__ == "__main__":
    main()
