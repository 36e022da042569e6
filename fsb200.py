"""Import alias: `import fsb200` loads the package directory
`finetoolsflexstructures.jl_b200/` (its name contains a dot, so it cannot be imported by
name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "finetoolsflexstructures.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "fsb200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fsb200"] = _mod
_spec.loader.exec_module(_mod)
