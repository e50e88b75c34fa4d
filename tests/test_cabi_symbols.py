"""CPU-side checks of the drop-in boundary: libgp2d.so builds, loads, and exports every
function include/gp2d.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gp2d.h")
LIB = os.path.join(ROOT, "2d-gp_b200", "libgp2d.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gp2d_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "2d-gp_b200", "build.py")])
    return ctypes.CDLL(LIB)


def test_header_declares_the_path():
    names = declared_functions()
    for must in ["gp2d_kernel_build", "gp2d_kdiag", "gp2d_kernel_grad", "gp2d_potrf", "gp2d_fit",
                 "gp2d_predict", "gp2d_lml_grad", "gp2d_fit_predict_host"]:
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), "libgp2d.so does not export %s" % name


def test_binding_table_matches_header(lib):
    import gp2d_b200
    from gp2d_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()
    assert _lib.lib.gp2d_version() == 100
    assert _lib.lib.gp2d_error_string(-3).decode() == "invalid argument 3"
    assert _lib.lib.gp2d_error_string(7).decode().startswith("matrix not positive definite")


def test_workspace_queries_and_argument_validation(lib):
    from gp2d_b200._lib import lib as L
    # pure host arithmetic: no device needed
    n400 = L.gp2d_fit_workspace_bytes(400)
    assert n400 >= 2 * 896 * 896 * 8           # A and Z padded to 896
    assert L.gp2d_fit_workspace_bytes(0) == 0
    assert L.gp2d_potrf_workspace_bytes(1000) >= 2 * 1024 * 1024 * 8
    assert L.gp2d_kernel_build(None, 4, None, 4, 1.0, 1.0, 0.5, 0.0, None, 8, None) == -1
    assert L.gp2d_predict(None, 4, 1.0, 1.0, 0.5, None, 1, 1, 0.0, None, None, None, 0, None) == -1


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    import gp2d_b200 as gp
    with pytest.raises(RuntimeError):
        gp.kernel_K(np.zeros((3, 2)), None, 1.0, 1.0, 0.5)
    with pytest.raises(RuntimeError):
        gp.HelmholtzGP(np.zeros((3, 2)), np.zeros(6), 1.0, 1.0, 0.5, 0.1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "2d-gp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def _build_c_example(tmp_path):
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "2d-gp_b200")
    exe = str(tmp_path / "host_example")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", "host_example.c"), "-L" + libdir, "-lgp2d",
                    "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    return exe


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """include/gp2d.h must be usable from C (the boundary is a C ABI): the plain-C example of
    INTEGRATION.md route C compiles with -std=c99 -Wall -Werror and links against libgp2d.so."""
    assert os.path.isfile(_build_c_example(tmp_path))


@pytest.mark.gpu
def test_c_example_runs(tmp_path):
    import subprocess
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = subprocess.run([_build_c_example(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "LML" in out.stdout


def test_argument_validation_of_every_family_returns_before_any_device_work():
    """Each entry point checks its arguments first and answers -(index of the bad argument) without
    touching the device: callable on a box without a GPU.  One probe per family and per kind of
    mistake (null pointer, size out of range, bad hyper-parameter, undersized workspace)."""
    import numpy as np
    from gp2d_b200._lib import lib as L
    dummy = np.zeros(16)
    p = dummy.ctypes.data                      # a non-null host address: never dereferenced by these calls
    ty = np.array([0, 1], dtype=np.int32)
    pr = np.array([[1.0, 1.0, 1.0, 1.0], [1.0, 1.0, 2.0, 2.0]])
    bad_pr = np.array([[1.0, 1.0, -1.0, 1.0], [1.0, 1.0, 2.0, 2.0]])
    var, ls = np.array([1.0]), np.array([[1.0, 2.0, 3.0]])
    # Helmholtz
    assert L.gp2d_fit(None, 4, p, 1.0, 1.0, 0.5, 0.1, 0.0, p, 1 << 30, None, None, None, None) == -1
    assert L.gp2d_fit(p, 0, p, 1.0, 1.0, 0.5, 0.1, 0.0, p, 1 << 30, None, None, None, None) == -2
    assert L.gp2d_fit(p, 100000, p, 1.0, 1.0, 0.5, 0.1, 0.0, p, 1 << 30, None, None, None, None) == -2     # 2N > 131072
    assert L.gp2d_fit(p, 4, p, -1.0, 1.0, 0.5, 0.1, 0.0, p, 1 << 30, None, None, None, None) == -4
    assert L.gp2d_fit(p, 4, p, 1.0, 1.0, 1.5, 0.1, 0.0, p, 1 << 30, None, None, None, None) == -4
    assert L.gp2d_fit(p, 4, p, 1.0, 1.0, 0.5, -0.1, 0.0, p, 1 << 30, None, None, None, None) == -7
    assert L.gp2d_fit(p, 4, p, 1.0, 1.0, 0.5, 0.1, 0.0, p, 16, None, None, None, None) == -10               # workspace too small
    assert L.gp2d_lml_grad(p, 4, p, 1.0, 1.0, 0.5, 0.1, 0.0, 0, p, 1 << 30, None, None, None) == -12
    assert L.gp2d_kernel_grad(p, 4, None, 5, 1.0, 1.0, 0.5, 0, p, 10, p, 1 << 20, p, None) == -4            # X2 == NULL needs M == N
    assert L.gp2d_potrf(p, 4, 3, p, 1 << 30, p, None) == -3                                                 # lda < n
    assert L.gp2d_dgemm(0, 0, 100, 128, 16, 1.0, p, 16, p, 128, 0.0, p, 128, None) == -3                     # M not a multiple of 128
    # space-time product
    assert L.gp2d_st_fit(p, 4, p, 1.0, 1.0, 0.5, 0.0, 1.0, 0.1, 0.0, p, 1 << 30, None, None, None, None) == -7   # tvar must be > 0
    assert L.gp2d_st_predict(p, 4, 1.0, 1.0, 0.5, 1.0, 1.0, None, 3, 3, 0.0, p, p, p, 1 << 30, None) == -8
    # sum of terms
    assert L.gp2d_hsum_fit(p, 4, 3, p, 2, ty.ctypes.data, bad_pr.ctypes.data, 0.1, 0.0, p, 1 << 30, None, None, None, None) == -5
    assert L.gp2d_hsum_fit(p, 4, 4, p, 2, ty.ctypes.data, pr.ctypes.data, 0.1, 0.0, p, 1 << 30, None, None, None, None) == -5    # ldx is 2 or 3
    assert L.gp2d_hsum_fit(p, 4, 3, p, 9, ty.ctypes.data, pr.ctypes.data, 0.1, 0.0, p, 1 << 30, None, None, None, None) == -5    # at most 8 terms
    assert L.gp2d_hsum_fit(p, 4, 3, p, 2, ty.ctypes.data, pr.ctypes.data, 0.1, 0.0, p, 16, None, None, None, None) == -11
    assert L.gp2d_hsum_predict(p, 4, 3, 2, ty.ctypes.data, pr.ctypes.data, p, 5, 4, 0.0, p, p, p, 1 << 30, None) == -9           # out_stride < M
    assert L.gp2d_hsum_fit_workspace_bytes(4, 3, 0) == 0 and L.gp2d_hsum_fit_workspace_bytes(4, 3, 2) > 0
    assert L.gp2d_hsum_kernel_grad_workspace_bytes(100, 100, 8) == 8 * L.gp2d_hsum_kernel_grad_workspace_bytes(100, 100, 1)
    # scalar RBF sums
    assert L.gp2d_rbf_fit(p, 4, 5, p, 1, var.ctypes.data, ls.ctypes.data, 0.1, 0.0, p, 1 << 30, None, None, None, None) == -5  # D <= 4
    assert L.gp2d_rbf_fit(p, 4, 3, p, 1, var.ctypes.data, ls.ctypes.data, 0.1, -1.0, p, 1 << 30, None, None, None, None) == -9
    assert L.gp2d_rbf_predict(p, 4, 3, 1, var.ctypes.data, ls.ctypes.data, p, 3, 0.0, None, p, p, 1 << 30, None) == -10
    assert L.gp2d_rbf_fit_workspace_bytes(200000, 3) == 0
    # M == 0 is a valid no-op everywhere
    assert L.gp2d_predict(p, 4, 1.0, 1.0, 0.5, None, 0, 0, 0.0, None, None, None, 0, None) == 0
    assert L.gp2d_hsum_predict(p, 4, 3, 2, ty.ctypes.data, pr.ctypes.data, None, 0, 0, 0.0, None, None, None, 0, None) == 0
    assert L.gp2d_error_string(-1005).decode().startswith("CUDA error")
