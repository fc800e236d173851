"""Import alias: the package directory is ``gan-image-captioning_b200/`` (not a valid Python identifier),
so ``import gic_b200`` loads it under this name.  ``gic_b200.generator`` etc. then import normally."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gan-image-captioning_b200")
_spec = importlib.util.spec_from_file_location("gic_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gic_b200"] = _mod
_spec.loader.exec_module(_mod)
