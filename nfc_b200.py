"""Import alias: `import nfc_b200` loads the package in `neuralfreeconvection.jl_b200/` (a dotted directory
name cannot be imported with a plain `import` statement)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "neuralfreeconvection.jl_b200")
_name = "neuralfreeconvection_jl_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_name]
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
__all__ = list(_pkg.__all__)
