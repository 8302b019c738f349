import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


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
