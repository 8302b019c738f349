"""Drop-in ``import MTM`` shim: resolves to the B200-native package in
``multitemplatematching-python_b200/`` (the directory name is not a valid Python
identifier, so it is loaded by path under the module name ``mtm_b200``)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multitemplatematching-python_b200")


def _load():
    if "mtm_b200" in sys.modules:
        return sys.modules["mtm_b200"]
    spec = importlib.util.spec_from_file_location("mtm_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["mtm_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


_impl = _load()
from mtm_b200 import (TRANSFORMS, BBox, TemplateTuple, __version__, computeScoreMap, drawBoxesOnGray,  # noqa: E402,F401
                      drawBoxesOnRGB, expandTemplates, findMatches, matchTemplates, matchTemplatesAugmented,
                      matchTemplatesBatch, matchTemplatesPyramid)
# as in the reference (MTM/__init__.py:13): importing the submodule first, then rebinding the
# name, leaves ``MTM.NMS`` bound to the FUNCTION while ``from MTM.NMS import NMS`` also works
from .NMS import NMS, Hit  # noqa: E402,F401

__all__ = ["NMS"]
