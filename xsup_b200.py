"""Importable alias for the package directory `x-as-supervision_b200/` (a hyphen is not a valid
identifier): `import xsup_b200` returns that package."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("x-as-supervision_b200")
