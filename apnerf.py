"""Import alias: ``import apnerf`` gives the package whose directory name
(``active-perception-using-neural-radiance-fields_b200``) is not a Python identifier."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("active-perception-using-neural-radiance-fields_b200")
sys.modules[__name__] = _pkg
