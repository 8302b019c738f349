import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")
    if os.environ.get("MTM_B200_EMULATE") == "1":
        # TEST INFRASTRUCTURE, opt-in: run the -m gpu tests WITHOUT a GPU against the host build of the whole library
        # (tests/emu_library.py: every kernel on the CPU emulation, the tcgen05 kernels on the functional model; slow --
        # select small cases with -k).  tests/test_library_emulation.py runs a curated selection this way.
        import tempfile
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import emu_library
        import MTM  # noqa: F401  (creates the mtm_b200 module alias)
        from mtm_b200 import _native
        _native.LIB_PATH = emu_library.build(tempfile.mkdtemp(prefix="mtm_emu_"))
        _native._lib = None
        os.environ["MTM_B200_EMULATED_LIB"] = _native.LIB_PATH      # for tests that start their own interpreter (test_gpu_knobs.py)


@pytest.fixture(scope="session")
def mtm():
    """The product API; loading it on a GPU-less box must fail loudly, not fall back."""
    import MTM
    return MTM


@pytest.fixture(scope="session")
def golden():
    import json
    from oracle import golden_cases as gc
    with open(os.path.join(gc.GOLDEN_DIR, "ref_outputs.json")) as f:
        return json.load(f)
