"""Import alias: `import b200dit` loads the package in ./omnihuman-1-hack_b200/ (whose directory
name is not a valid Python identifier)."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "omnihuman-1-hack_b200")
_spec = importlib.util.spec_from_file_location("b200dit", os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200dit"] = _mod
_spec.loader.exec_module(_mod)
