"""CPU oracle for the 2D-GP hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product path
(``2d-gp_b200/``) never imports it and has no CPU fallback.

Pinning status: the reference ships no tests, assertions or stored outputs
(SURVEY.md §4, §8c).  The kernel-value / layout / gradient-formula parts of the
oracle are pinned against *verbatim line slices of the reference itself* run in
the build container (``oracle/ref_slices.py`` -> ``tests/golden/*.npz``).  The
fit / predict / log-likelihood algebra lives in un-vendored third-party code
(GPy, scikit-learn, LAPACK) and is pinned against the reference's own numpy
formulation (GP_laser.py:113-140,177-185: explicit inverse) on simulTracks.pkl
and against live scikit-learn for scalar kernels; for the GPy boundary itself:
"parity unpinned".  Two kernel families restate formulas for which the reference
has no runnable code at all -- the space-time product Kt * nonDivK (dead code
upstream) and the myKernel2 divFreeK / curlFreeK sums (module missing upstream):
"parity unpinned" for both, except on their isotropic subspaces, where they must
reproduce the pinned Helmholtz kernel (tests/test_oracle_golden.py).
"""
from .gp_oracle import *  # noqa: F401,F403
