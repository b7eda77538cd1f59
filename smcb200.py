"""Importable alias for the package directory (whose name, fixed by the project layout, has hyphens)."""
import importlib
import os
import sys

_NAME = "python-based-sequential-monte-carlo-method-with-likelihood-tempering_b200"
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module(_NAME)
sys.modules[__name__] = _pkg
