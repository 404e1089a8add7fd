"""Importable alias for the package directory `datafusion-ballista_b200/` (a hyphen is not a valid
Python identifier): `import ballista_b200` loads that directory as a package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "datafusion-ballista_b200")
_spec = importlib.util.spec_from_file_location("ballista_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ballista_b200"] = _mod
_spec.loader.exec_module(_mod)
