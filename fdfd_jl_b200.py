"""Import shim: the product package lives in the directory `fdfd.jl_b200/` (named after the reference,
FDFD.jl); a dot is not importable, so this module loads it under the name `fdfd_jl_b200`.

    import fdfd_jl_b200 as fdfd
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "fdfd.jl_b200")
_spec = _ilu.spec_from_file_location("fdfd_jl_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["fdfd_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
