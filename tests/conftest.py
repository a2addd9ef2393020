import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_built():
    """Builds the oracle's C port (and oracle/_ref when the reference tree is mounted)."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    return True


@pytest.fixture(scope="session")
def fixtures():
    from tests._cases import load_fixtures
    return load_fixtures()


@pytest.fixture(scope="session")
def golden():
    from tests._cases import load_golden
    return load_golden()
