"""Import shim: makes the package directory ``starformationhistories.jl_b200/`` (whose name contains a dot)
importable as ``sfh_b200`` -- ``import sfh_b200`` / ``from sfh_b200 import fitting``."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "starformationhistories.jl_b200")
_spec = _u.spec_from_file_location("sfh_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["sfh_b200"] = _mod
_spec.loader.exec_module(_mod)
