"""Import alias: ``import pfc_b200`` loads the package living in ``pressurefieldcontact.jl_b200/``
(a directory name with a dot cannot be imported by name)."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pressurefieldcontact.jl_b200")
_spec = importlib.util.spec_from_file_location("pfc_b200", os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pfc_b200"] = _mod
_spec.loader.exec_module(_mod)
