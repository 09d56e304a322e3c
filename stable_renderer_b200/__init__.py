"""Importable alias of the `stable-renderer_b200/` source directory (a hyphen is not a valid module name).

All code lives in `../stable-renderer_b200/`; this stub only points the package search path there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "stable-renderer_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
