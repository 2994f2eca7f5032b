"""Import alias: ``import hy3dgeo`` == the package in ``hunyuan3d-2_b200/`` (whose
directory name is not a valid Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("hunyuan3d-2_b200")
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("hunyuan3d-2_b200."):
        sys.modules["hy3dgeo." + _name.split(".", 1)[1]] = _mod
sys.modules[__name__] = _pkg
