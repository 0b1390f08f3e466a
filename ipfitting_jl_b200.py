"""Import shim: the package directory is literally `ipfitting.jl_b200/` (a dot is
not importable), so this module exposes it as the package `ipfitting_jl_b200`."""
import os as _os

__package__ = __name__
__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "ipfitting.jl_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
