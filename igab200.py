"""Import shim: the package directory is named `dune-iga_b200` (not a legal Python identifier), so this module
registers it as `dune_iga_b200` and re-exports it.  `import igab200` then `from dune_iga_b200 import ...`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dune-iga_b200")
if "dune_iga_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location("dune_iga_b200", os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["dune_iga_b200"] = _mod
    _spec.loader.exec_module(_mod)
pkg = sys.modules["dune_iga_b200"]
