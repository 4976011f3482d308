"""Import shim: the package directory is ``laser-polio_b200/`` (the layout the build
brief names), which is not a valid Python identifier.  ``import laser_polio_b200``
resolves here and is redirected to that directory as a regular package."""

import importlib.util
import pathlib
import sys

_dir = pathlib.Path(__file__).resolve().parent / "laser-polio_b200"
_spec = importlib.util.spec_from_file_location(
    "laser_polio_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["laser_polio_b200"] = _mod
_spec.loader.exec_module(_mod)
