"""Import shim: the package sources live in `gf-nerf_b200/` (not a valid Python identifier)."""
import os as _os

_src = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gf-nerf_b200")
__path__.insert(0, _src)
with open(_os.path.join(_src, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_src, "__init__.py"), "exec"))
del _f
