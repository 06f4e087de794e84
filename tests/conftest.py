import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/libswirl_oracle.so), built on demand. Checker only."""
    path = os.path.join(ROOT, "oracle", "libswirl_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    import oracle_lib

    return oracle_lib.Oracle(path)


@pytest.fixture(scope="session")
def dev():
    """The product device (CUDA only; fails loudly when the library or a GPU is missing)."""
    import stark_backend_b200 as sb

    d = sb.B200Device(0)
    yield d
    d.close()
