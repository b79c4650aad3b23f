"""Imports the package directory `odr-dabmod_b200/` (a hyphen is not a valid
module name) as module `odr_dabmod_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    name = "odr_dabmod_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(_ROOT, "odr-dabmod_b200")
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
