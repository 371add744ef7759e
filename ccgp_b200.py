"""Import alias: the package directory is named after the reference repo
(`convex-combination-of-gaussian-processes_b200/`), which is not a valid Python
identifier; `import ccgp_b200` loads that directory as a package."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "convex-combination-of-gaussian-processes_b200")
_spec = importlib.util.spec_from_file_location(
    "ccgp_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ccgp_b200"] = _mod
_spec.loader.exec_module(_mod)
