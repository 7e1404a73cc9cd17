"""Import alias: the package directory is `manta-rs_b200/` (not a Python identifier).

`import manta_rs_b200` loads that directory as a regular package under this name.
"""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "manta-rs_b200")
_spec = importlib.util.spec_from_file_location(
    "manta_rs_b200", os.path.join(_root, "__init__.py"), submodule_search_locations=[_root]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["manta_rs_b200"] = _mod
_spec.loader.exec_module(_mod)
