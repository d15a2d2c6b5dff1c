import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_api():
    from oracle import oracle_ffi
    oracle_ffi.build()
    return oracle_ffi.load()


@pytest.fixture(scope="session")
def gpu_api():
    from pbrs_b200 import _ffi
    return _ffi.load()


@pytest.fixture(scope="session")
def hostsim_api():
    from tests import hostsim
    return hostsim.load()
