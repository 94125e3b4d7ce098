"""Import shim: exposes the package directory `3dsmc-bundle-adjustment_b200/`
(not a valid Python identifier) as the module `ba_b200`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "3dsmc-bundle-adjustment_b200")
_spec = importlib.util.spec_from_file_location("ba_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ba_b200"] = _mod
_spec.loader.exec_module(_mod)
