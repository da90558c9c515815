import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/, test infrastructure).  Built on demand with gcc."""
    from oracle import oracle as orc
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def engine():
    """The product engine through its C-ABI (ctypes).  GPU tests only."""
    import trgt_b200
    eng = trgt_b200.Engine(device=0)
    yield eng
    eng.close()
