"""Import alias: `scvx_b200` is the package whose sources live in
`dynamic-programming-multiagent-trajectory-optimiziation_b200/` (not a valid identifier)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "dynamic-programming-multiagent-trajectory-optimiziation_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
