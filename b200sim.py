"""Import alias: ``import b200sim`` loads the package in ``3d-spatial-sim-for-boid-and-nbody_b200/``
(a directory name that is not a Python identifier) under the module name ``b200sim``."""
import importlib.util as _u
import os as _os
import sys as _sys

_PKG = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "3d-spatial-sim-for-boid-and-nbody_b200")
_spec = _u.spec_from_file_location("b200sim", _os.path.join(_PKG, "__init__.py"), submodule_search_locations=[_PKG])
_mod = _u.module_from_spec(_spec)
_sys.modules["b200sim"] = _mod
_spec.loader.exec_module(_mod)
