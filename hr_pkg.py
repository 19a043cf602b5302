"""Loader for the package directory `mpv-frame-interpolator_b200/`.

The directory name is fixed by the project layout and contains a hyphen, so it cannot be
imported with a plain `import`; everything in the repo (tests, bench.py, __graft_entry__.py)
gets it through `hr_pkg.load()` under the module name `hopperrender_b200`.
"""
import importlib.util
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent
PKG_DIR = ROOT / "mpv-frame-interpolator_b200"
_NAME = "hopperrender_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, PKG_DIR / "__init__.py", submodule_search_locations=[str(PKG_DIR)]
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
