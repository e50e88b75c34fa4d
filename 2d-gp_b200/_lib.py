"""ctypes binding of libgp2d.so (the C ABI declared in include/gp2d.h).

There is no fallback: if the shared library is missing or a symbol is absent the import
raises.  Build it with ``python 2d-gp_b200/build.py`` or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgp2d.so")

c_dp = C.c_void_p      # device / host double*
c_ip = C.c_void_p      # int*
c_st = C.c_void_p      # cudaStream_t

# name -> (restype, argtypes); must list every symbol of include/gp2d.h
SIGNATURES = {
    "gp2d_version": (C.c_int, []),
    "gp2d_error_string": (C.c_char_p, [C.c_int]),
    "gp2d_kernel_build": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, C.c_double, C.c_double, C.c_double,
                                    C.c_double, c_dp, C.c_int64, c_st]),
    "gp2d_kdiag": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, c_dp, c_st]),
    "gp2d_kernel_grad_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gp2d_kernel_grad": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                   c_dp, C.c_int64, C.c_void_p, C.c_size_t, c_dp, c_st]),
    "gp2d_potrf_workspace_bytes": (C.c_size_t, [C.c_int]),
    "gp2d_potrf": (C.c_int, [c_dp, C.c_int, C.c_int64, C.c_void_p, C.c_size_t, c_ip, c_st]),
    "gp2d_spd_inverse_workspace_bytes": (C.c_size_t, [C.c_int]),
    "gp2d_spd_inverse": (C.c_int, [c_dp, C.c_int, C.c_int64, C.c_void_p, C.c_size_t, c_ip, c_st]),
    "gp2d_dgemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, c_dp, C.c_int64,
                             c_dp, C.c_int64, C.c_double, c_dp, C.c_int64, c_st]),
    "gp2d_fit_workspace_bytes": (C.c_size_t, [C.c_int]),
    "gp2d_fit": (C.c_int, [c_dp, C.c_int, c_dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                           C.c_void_p, C.c_size_t, c_dp, c_dp, c_ip, c_st]),
    "gp2d_fit_predict_state": (C.c_int, [C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "gp2d_predict_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gp2d_predict": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, c_dp, C.c_int,
                               C.c_int64, C.c_double, c_dp, c_dp, C.c_void_p, C.c_size_t, c_st]),
    "gp2d_lml_grad": (C.c_int, [c_dp, C.c_int, c_dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                C.c_int, C.c_void_p, C.c_size_t, c_dp, c_ip, c_st]),
    "gp2d_st_kernel_build": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_double, c_dp, C.c_int64, c_st]),
    "gp2d_st_kernel_grad": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                      C.c_double, c_dp, C.c_int64, C.c_void_p, C.c_size_t, c_dp, c_st]),
    "gp2d_st_fit_workspace_bytes": (C.c_size_t, [C.c_int]),
    "gp2d_st_fit_predict_state": (C.c_int, [C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "gp2d_st_fit": (C.c_int, [c_dp, C.c_int, c_dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                              C.c_double, C.c_double, C.c_void_p, C.c_size_t, c_dp, c_dp, c_ip, c_st]),
    "gp2d_st_predict": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                  c_dp, C.c_int, C.c_int64, C.c_double, c_dp, c_dp, C.c_void_p, C.c_size_t, c_st]),
    "gp2d_st_lml_grad": (C.c_int, [c_dp, C.c_int, c_dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_double, C.c_double, C.c_void_p, C.c_size_t, c_dp, c_ip, c_st]),
    "gp2d_rbf_kernel_build": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_dp, C.c_double, c_dp,
                                        C.c_int64, c_st]),
    "gp2d_rbf_kernel_grad_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gp2d_rbf_kernel_grad": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, C.c_int64,
                                       C.c_void_p, C.c_size_t, c_dp, c_st]),
    "gp2d_rbf_fit_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gp2d_rbf_fit_predict_state": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "gp2d_rbf_fit": (C.c_int, [c_dp, C.c_int, C.c_int, c_dp, C.c_int, c_dp, c_dp, C.c_double, C.c_double,
                               C.c_void_p, C.c_size_t, c_dp, c_dp, c_ip, c_st]),
    "gp2d_rbf_predict_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gp2d_rbf_predict": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, C.c_int, C.c_double,
                                   c_dp, c_dp, C.c_void_p, C.c_size_t, c_st]),
    "gp2d_rbf_lml_grad": (C.c_int, [c_dp, C.c_int, C.c_int, c_dp, C.c_int, c_dp, c_dp, C.c_double, C.c_double,
                                    C.c_void_p, C.c_size_t, c_dp, c_ip, c_st]),
    "gp2d_hsum_kernel_build": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, C.c_int, C.c_int, c_ip, c_dp, C.c_double, c_dp,
                                         C.c_int64, c_st]),
    "gp2d_hsum_kdiag": (C.c_int, [C.c_int, C.c_int, C.c_int, c_ip, c_dp, c_dp, c_st]),
    "gp2d_hsum_kernel_grad_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gp2d_hsum_kernel_grad": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, C.c_int, C.c_int, c_ip, c_dp, c_dp, C.c_int64,
                                        C.c_void_p, C.c_size_t, c_dp, c_st]),
    "gp2d_hsum_fit_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gp2d_hsum_fit_predict_state": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "gp2d_hsum_fit": (C.c_int, [c_dp, C.c_int, C.c_int, c_dp, C.c_int, c_ip, c_dp, C.c_double, C.c_double,
                                C.c_void_p, C.c_size_t, c_dp, c_dp, c_ip, c_st]),
    "gp2d_hsum_predict": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_ip, c_dp, c_dp, C.c_int, C.c_int64,
                                    C.c_double, c_dp, c_dp, C.c_void_p, C.c_size_t, c_st]),
    "gp2d_hsum_lml_grad": (C.c_int, [c_dp, C.c_int, C.c_int, c_dp, C.c_int, c_ip, c_dp, C.c_double, C.c_double,
                                     C.c_void_p, C.c_size_t, c_dp, c_ip, c_st]),
    "gp2d_fit_predict_host": (C.c_int, [c_dp, C.c_int, c_dp, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_double, c_dp, C.c_int, C.c_int, c_dp, c_dp, c_dp]),
    "gp2d_host_release": (None, []),
    "gp2d_set_option": (C.c_int, [C.c_int, C.c_double]),
    "gp2d_fit_batched_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gp2d_fit_batched": (C.c_int, [c_dp, C.c_int64, C.c_int, c_dp, C.c_int64, C.c_int, c_dp, C.c_double, C.c_void_p,
                                   C.c_size_t, c_dp, c_dp, c_ip, c_st]),
    "gp2d_lml_grad_batched": (C.c_int, [c_dp, C.c_int64, C.c_int, c_dp, C.c_int64, C.c_int, c_dp, C.c_double, C.c_int,
                                        C.c_void_p, C.c_size_t, c_dp, c_ip, c_st]),
}


GP2D_OPT_PREDICT_I8 = 1


class Gp2dError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "gp2d: %s not found -- the CUDA extension is required (no CPU fallback). "
            "Build it with `python 2d-gp_b200/build.py`." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str) -> None:
    if rc < 0:
        raise Gp2dError("%s failed (%d): %s" % (what, rc, lib.gp2d_error_string(rc).decode()))
