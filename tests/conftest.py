import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def ref():
    """The CPU oracle (oracle/pyoracle.py) -- the checker, never the thing under test in -m gpu."""
    import pyoracle

    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def fixtures():
    from helpers import load_fixtures

    return load_fixtures()
