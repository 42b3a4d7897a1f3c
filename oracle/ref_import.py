"""Import the unmodified reference (``/root/reference`` when present, else the staged copy in ``oracle/_ref``).

Test / benchmark infrastructure only.  Shims: empty ``matplotlib`` modules (alignment.py:23 imports
pylab at module level; nothing on the decoder path draws), ``vocoders/`` on ``sys.path`` (decoders.py:30-31).
"""
from __future__ import annotations

import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("RADMMM_REFERENCE", "/root/reference"), os.path.join(HERE, "_ref")]


def reference_root():
    for root in _CANDIDATES:
        if root and os.path.exists(os.path.join(root, "decoders.py")):
            return root
    return None


def available() -> bool:
    return reference_root() is not None


def import_reference():
    """Returns (root, modules) with modules = {'decoders', 'common', 'loss', 'radam'}; raises ImportError if absent."""
    root = reference_root()
    if root is None:
        raise ImportError("the reference is neither at /root/reference nor staged under oracle/_ref "
                          "(run `python -m oracle.stage_ref` in the build container)")
    for p in (os.path.join(root, "vocoders"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ("matplotlib", "matplotlib.pylab"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]
    warnings.filterwarnings("ignore")
    import importlib
    mods = {n: importlib.import_module(n) for n in ("common", "decoders", "loss", "radam")}
    return root, mods
