"""Importable alias of the package directory `se3-equi-graph-registration_b200/` (a hyphen is
not a legal Python identifier).  All code lives there; this shim only redirects the package
search path and executes that directory's __init__."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "se3-equi-graph-registration_b200")
__path__[:] = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
