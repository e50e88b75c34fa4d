"""Import shim: the package directory is named ``2d-gp_b200`` (not a Python identifier),
so ``import gp2d_b200`` resolves here, turns this module into a package whose ``__path__``
is that directory, and executes the package ``__init__`` (sub-modules import normally)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "2d-gp_b200")]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
