"""Import shim: makes the literal directory `grape.jl_b200/` importable as `grape.jl_b200`.

`import grape.jl_b200` resolves the package `grape` (this file) and then the
submodule `jl_b200`; we register the sibling directory `../grape.jl_b200` under
that name so the on-disk layout can keep the reference-derived name.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "grape.jl_b200")
_name = __name__ + ".jl_b200"
if _name not in _sys.modules:
    _spec = _ilu.spec_from_file_location(
        _name, _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
    jl_b200 = _ilu.module_from_spec(_spec)
    _sys.modules[_name] = jl_b200
    _spec.loader.exec_module(jl_b200)
else:  # pragma: no cover
    jl_b200 = _sys.modules[_name]
