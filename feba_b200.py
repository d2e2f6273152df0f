"""Import shim: ``import feba_b200`` loads the package directory
``fish-eye_bundle_adjustment_b200/`` (its name is not a valid Python identifier)."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "fish-eye_bundle_adjustment_b200")
_spec = importlib.util.spec_from_file_location(
    "feba_b200", os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["feba_b200"] = _mod
_spec.loader.exec_module(_mod)
