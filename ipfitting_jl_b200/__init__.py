"""Import shim: the package directory is literally `ipfitting.jl_b200/` (a dot is not
importable), so this stub package redirects `ipfitting_jl_b200` to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ipfitting.jl_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
