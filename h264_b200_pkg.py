"""Import helper: registers the package directory `h.264_b200/` (not a valid identifier) as module `h264_b200`."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "h.264_b200")


def load():
    if "h264_b200" in sys.modules:
        return sys.modules["h264_b200"]
    spec = importlib.util.spec_from_file_location("h264_b200", os.path.join(_DIR, "__init__.py"),
                                                  submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["h264_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = load()
