"""Import the UNMODIFIED reference package (TEST INFRASTRUCTURE, build container only).

``/root/reference`` exists only in the build container.  ``import MTM`` there
fails because scikit-image is not installed (``MTM/__init__.py:11``); the only
symbol it needs is ``skimage.feature.peak_local_max``, so a stand-in module that
exposes ``oracle.peaks.peak_local_max`` is registered in ``sys.modules`` before
the import.  Nothing in the reference tree is modified or copied.

Never call this from ``-m gpu`` tests, ``smoke()`` or ``bench.py``: the GPU box
has no /root/reference.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = "/root/reference"


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "MTM", "__init__.py"))


def load(name="_mtm_reference"):
    """Returns the reference ``MTM`` package loaded under a private module name."""
    if name in sys.modules:
        return sys.modules[name]
    if not available():
        raise RuntimeError("reference tree not present (only in the build container)")
    if "skimage" not in sys.modules:
        from . import peaks
        sk = types.ModuleType("skimage")
        feat = types.ModuleType("skimage.feature")
        feat.peak_local_max = peaks.peak_local_max
        sk.feature = feat
        sys.modules["skimage"] = sk
        sys.modules["skimage.feature"] = feat
    pkg_dir = os.path.join(REF_ROOT, "MTM")
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
