"""``from MTM.NMS import NMS`` compatibility (reference: MTM/NMS.py)."""
from mtm_b200.api import NMS  # noqa: F401

Hit = tuple
