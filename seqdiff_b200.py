"""Importable alias for the package directory `e3-invaraint-diffusion-model_b200/` (its name is not a
valid Python identifier):  `import seqdiff_b200 as sd`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("e3-invaraint-diffusion-model_b200")
sys.modules[__name__] = _pkg
