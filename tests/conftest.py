import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cam0():
    """The bundled cam0 data set as frozen by tests/golden/make_golden.py (numeric arrays)."""
    from tests.golden import load_cam0
    return load_cam0
