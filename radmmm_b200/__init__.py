"""Importable name of the product package.

The sources live in ``rad-mmm_b200/`` (the repository's layout contract; a hyphen cannot appear in a Python package name).
This package's search path IS that directory, so ``radmmm_b200.decoders`` is ``rad-mmm_b200/decoders.py`` through the
ordinary import machinery -- nothing is exec'd or copied.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rad-mmm_b200")]
__version__ = "0.2.0"
