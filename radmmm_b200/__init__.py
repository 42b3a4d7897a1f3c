"""Import shim: the product package lives in ``rad-mmm_b200/`` (a name Python cannot import directly).

``import radmmm_b200`` executes ``rad-mmm_b200/__init__.py`` under this importable name and points the
package search path there, so ``radmmm_b200.decoders`` is ``rad-mmm_b200/decoders.py``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rad-mmm_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
