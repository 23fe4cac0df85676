"""Import shim: exposes the package in ``carefree-learn_b200/`` (not a valid identifier) as ``cflearn_b200``."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "carefree-learn_b200")
_spec = importlib.util.spec_from_file_location(
    "cflearn_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
)
_module = importlib.util.module_from_spec(_spec)
sys.modules["cflearn_b200"] = _module
_spec.loader.exec_module(_module)
