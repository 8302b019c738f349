"""mtm_b200 -- B200-native drop-in for the MTM 2.0.1 hot path.

``matchTemplates`` / ``findMatches`` / ``computeScoreMap`` / ``NMS`` keep the
reference's signatures (MTM/__init__.py:56,95,247; MTM/NMS.py:20) and run on
hand-written sm_100a kernels behind the C ABI of ``libmtm_b200.so``.
"""
from .api import NMS, computeScoreMap, findMatches, matchTemplates, matchTemplatesBatch
from .augment import TRANSFORMS, expandTemplates, matchTemplatesAugmented, matchTemplatesPyramid
from .draw import drawBoxesOnGray, drawBoxesOnRGB
from ._native import Context, default_context

__version__ = "2.0.1"          # API level of the reference this mirrors (MTM/version.py:5)
Hit = tuple                    # (label, (x, y, width, height), score), MTM/NMS.py:18
BBox = tuple
TemplateTuple = tuple

__all__ = ["NMS"]              # as in MTM/__init__.py:16
