"""Import helper: the package directory is `h-numo_b200/` (hyphen), exposed as module `hnumo_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def _load():
    if "hnumo_b200" in sys.modules:
        return sys.modules["hnumo_b200"]
    pkg = os.path.join(_ROOT, "h-numo_b200")
    spec = importlib.util.spec_from_file_location("hnumo_b200", os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["hnumo_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


hnumo_b200 = _load()
