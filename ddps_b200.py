"""Import shim: ``import ddps_b200`` loads the package directory ``driftdiffusionpoissonsystems.jl_b200/``
(its literal name contains a dot and cannot be written in an import statement)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "driftdiffusionpoissonsystems.jl_b200")
_spec = importlib.util.spec_from_file_location("ddps_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ddps_b200"] = _mod
_spec.loader.exec_module(_mod)
