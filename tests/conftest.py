import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fdem():
    """The product package with the native library loaded.  GPU tests must run the CUDA path:
    a missing library is an error, never a skip."""
    import fastdem_b200
    fastdem_b200.load_library()
    return fastdem_b200
