"""Import shim: the package directory is named `alps-looper_b200` (not a Python identifier), so
load it by path and re-export it as the module `looper_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "alps-looper_b200")
_spec = importlib.util.spec_from_file_location("alps_looper_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules.setdefault("alps_looper_b200", _mod)
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
