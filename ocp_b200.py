"""Import alias: ``import ocp_b200`` loads the package directory whose (contract-mandated)
name contains hyphens and therefore cannot be written in an import statement."""
import importlib.util
import os
import sys

_DIR = os.path.join(
    os.path.dirname(os.path.abspath(__file__)),
    "optimal-control-of-a-coupled-navier-stokes-ode-system-for-reconstruction-of-ocean-currents_b200",
)
_spec = importlib.util.spec_from_file_location(
    "ocp_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ocp_b200"] = _mod
_spec.loader.exec_module(_mod)
