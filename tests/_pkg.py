"""Imports the package directory `unet-studio_b200/` (hyphenated, so not importable by name) as `unet_studio_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    if "unet_studio_b200" in sys.modules:
        return sys.modules["unet_studio_b200"]
    path = os.path.join(ROOT, "unet-studio_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("unet_studio_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["unet_studio_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
