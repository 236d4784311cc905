"""Import alias: the package directory is named after the reference
(`gradientdomain-mitsuba_b200/`, not a valid Python identifier), so this module
registers it under the importable name ``gdb200``."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "gradientdomain-mitsuba_b200")
_spec = importlib.util.spec_from_file_location(
    "gdb200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gdb200"] = _mod
_spec.loader.exec_module(_mod)
