"""Import alias: the product package lives in the directory ``lowrankintegrators.jl_b200/`` (the name the
build contract fixes), which is not a valid dotted Python name on its own.  Importing
``lowrankintegrators.jl_b200`` resolves to that directory through this tiny namespace shim."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "lowrankintegrators.jl_b200")
_spec = _ilu.spec_from_file_location("lowrankintegrators.jl_b200", _os.path.join(_pkg_dir, "__init__.py"),
                                     submodule_search_locations=[_pkg_dir])
jl_b200 = _ilu.module_from_spec(_spec)
_sys.modules["lowrankintegrators.jl_b200"] = jl_b200
_spec.loader.exec_module(jl_b200)
