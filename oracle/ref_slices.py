"""Run verbatim line slices of the reference under Python 3.  TEST INFRASTRUCTURE.

The reference modules cannot be imported whole here (py2 ``print``, ``import GPy``,
leading-zero literals), but the numerical functions are py3-clean.
This loader reads the line ranges *where they lie* under ``/root/reference`` (nothing
is copied into this repo), de-indents class methods, and ``exec``s them.  It is used
only by ``tests/golden/make_golden.py`` and by tests that skip when the reference tree
is absent (it does not exist on the GPU box).

The reference tree is untrusted content and this module EXECUTES parts of it, so it is opt-in: nothing
runs unless the environment has GP2D_EXEC_REFERENCE=1 (``available()`` is False otherwise and the one
test that uses live slices skips; ``tests/golden/make_golden.py`` asks for the flag explicitly).  The
committed fixtures under tests/golden/ keep pinning the oracle either way.
"""
from __future__ import annotations

import os

import numpy as np

REF = os.environ.get("GP2D_REFERENCE", "/root/reference")


def enabled() -> bool:
    return os.environ.get("GP2D_EXEC_REFERENCE") == "1"


def available() -> bool:
    return enabled() and os.path.isfile(os.path.join(REF, "GP_scripts.py"))


def _require():
    if not enabled():
        raise RuntimeError("oracle.ref_slices executes code from the reference tree: set GP2D_EXEC_REFERENCE=1 to allow it")


def _lines(fname, lo, hi):
    _require()
    with open(os.path.join(REF, fname), "r") as f:
        src = f.readlines()
    return "".join(src[lo - 1:hi])


def gp_scripts():
    """namespace with myKernel, getMean, getCov, nonDivK, compute_K, compute_Ks, sqExp, rbf
    (GP_scripts.py:1-3 imports + 6-142)."""
    ns = {}
    code = "import numpy as np\n" + _lines("GP_scripts.py", 6, 142)
    exec(compile(code, "GP_scripts.py[6:142]", "exec"), ns)
    return ns


class _P(float):
    """float with a settable ``.gradient`` (stands in for GPy's Param)."""
    gradient = None


class _Self:
    pass


def _method(fname, lo, hi, name):
    ns = {"np": np}
    # class methods sit at 4 spaces; some comment lines sit at column 0, so a common-prefix
    # dedent does not work: strip exactly one indent level where present.
    code = "".join(ln[4:] if ln.startswith("    ") else ln
                   for ln in _lines(fname, lo, hi).splitlines(keepends=True))
    exec(compile(code, "%s[%d:%d]" % (fname, lo, hi), "exec"), ns)
    return ns[name]


def mykernel_class(l_df, l_cf, ratio):
    """Stub ``self`` bound to myKernel.myKernel.K / Kdiag / update_gradients_full
    (myKernel.py:27-53, 55-57, 59-106)."""
    s = _Self()
    s.length_df, s.length_cf, s.ratio = _P(l_df), _P(l_cf), _P(ratio)
    K = _method("myKernel.py", 27, 53, "K")
    Kdiag = _method("myKernel.py", 55, 57, "Kdiag")
    upd = _method("myKernel.py", 59, 106, "update_gradients_full")
    s.K = lambda X, X2=None: K(s, X, X2)
    s.Kdiag = lambda X: Kdiag(s, X)
    s.update_gradients_full = lambda dL_dK, X, X2=None: upd(s, dL_dK, X, X2)
    return s


def nondivk_class(length):
    """nonDivK.K / Kdiag / update_gradients_full (myKernel.py:159-176,178-180,182-207)."""
    s = _Self()
    s.length = _P(length)
    K = _method("myKernel.py", 159, 176, "K")
    Kdiag = _method("myKernel.py", 178, 180, "Kdiag")
    upd = _method("myKernel.py", 182, 207, "update_gradients_full")
    s.K = lambda X, X2=None: K(s, X, X2)
    s.Kdiag = lambda X: Kdiag(s, X)
    s.update_gradients_full = lambda dL_dK, X, X2=None: upd(s, dL_dK, X, X2)
    return s


def nonrotk_class(length):
    """nonRotK.K / Kdiag (myKernel.py:255-271,273-275); its gradient method writes a
    non-existent attribute (myKernel.py:301) so only lines 277-299 (``dl``) are usable."""
    s = _Self()
    s.length = _P(length)
    K = _method("myKernel.py", 255, 271, "K")
    Kdiag = _method("myKernel.py", 273, 275, "Kdiag")
    s.K = lambda X, X2=None: K(s, X, X2)
    s.Kdiag = lambda X: Kdiag(s, X)
    s.length_cf = _P(length)     # target of the stray assignment at myKernel.py:301
    upd = _method("myKernel.py", 277, 301, "update_gradients_full")
    s.update_gradients_full = lambda dL_dK, X, X2=None: upd(s, dL_dK, X, X2)
    return s


class _Tracks(object):
    """Plain attribute container standing in for laser_class.interpolated_tracks."""


def load_simul_tracks():
    """simulTracks.pkl -> attribute container (py3: latin1).  A restricted unpickler: the pickle's
    laser_class.interpolated_tracks becomes a local plain container (nothing is imported from the
    reference directory), numpy arrays are rebuilt, every other global is refused.  Reading data is not
    executing reference code, so this does not need GP2D_EXEC_REFERENCE."""
    import pickle

    class _U(pickle.Unpickler):
        def find_class(self, module, name):
            if module == "laser_class" and name == "interpolated_tracks":
                return _Tracks
            if module.split(".")[0] == "numpy" and name in ("_reconstruct", "ndarray", "dtype", "scalar"):
                import importlib
                try:
                    return getattr(importlib.import_module(module), name)
                except (ImportError, AttributeError):
                    import numpy._core.multiarray as ma
                    return getattr(ma, name) if hasattr(ma, name) else getattr(np, name)
            if (module, name) in (("copy_reg", "_reconstructor"), ("copyreg", "_reconstructor")):
                import copyreg
                return copyreg._reconstructor
            if (module, name) in (("__builtin__", "object"), ("builtins", "object")):
                return object
            if module == "datetime" and name in ("datetime", "timedelta", "date"):
                import datetime
                return getattr(datetime, name)
            raise pickle.UnpicklingError("refusing to load %s.%s" % (module, name))

    with open(os.path.join(REF, "simulTracks.pkl"), "rb") as f:
        return _U(f, encoding="latin1").load()
