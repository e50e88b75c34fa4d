"""gp2d_b200 -- B200-native (sm_100a) Helmholtz Gaussian-process hot path of rafaelcgon/2D-GP.

Layers: csrc/ (CUDA kernels + C ABI, include/gp2d.h) -> _lib (ctypes) -> engine (device
tensors) -> reference call surface (myKernel, GP_scripts, GP_laser, krig, models).
Importing this package loads libgp2d.so and fails loudly when it is missing.
"""
from . import _lib                                           # noqa: F401  (loads the library)
from .engine import (HelmholtzGP, LinAlgError, kernel_K, kernel_Kdiag, kernel_grad_sums,   # noqa: F401
                     potrf, spd_inverse, matmul, fit_predict_host, as_dev, rbf_K, rbf_grad_sums, ScalarGP,
                     st_K, st_grad_sums, SpaceTimeGP, hsum_K, hsum_Kdiag, hsum_grad_sums, HelmholtzSumGP,
                     HelmholtzBatch, krig_snapshots, set_predict_i8)

__version__ = "0.1.0"
