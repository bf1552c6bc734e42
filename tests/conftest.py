import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import rfo
    rfo.build()
    rfo.load()
    return rfo


@pytest.fixture(scope="session")
def device():
    """One rf_ctx on cuda:0 for the whole session. Fails loudly if the library or GPU is missing."""
    import retrofire_b200 as rf
    dev = rf.Device(0)
    yield dev
    dev.close()
