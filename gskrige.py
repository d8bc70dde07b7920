"""Import shim: the package directory is named ``geostatssolvers.jl_b200`` (as the task fixes it),
which Python cannot import by name because of the dot. ``import gskrige`` executes this file,
which loads that directory as the package ``gskrige`` and replaces itself in ``sys.modules``."""
import importlib.util as _u
import sys as _sys
from pathlib import Path as _Path

_pkg = _Path(__file__).resolve().parent / "geostatssolvers.jl_b200"
_spec = _u.spec_from_file_location("gskrige", _pkg / "__init__.py", submodule_search_locations=[str(_pkg)])
_mod = _u.module_from_spec(_spec)
_sys.modules["gskrige"] = _mod
_spec.loader.exec_module(_mod)
