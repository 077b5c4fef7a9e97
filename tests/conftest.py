import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "big: full-size parity cases (minutes of host CPU for the oracle); part of -m gpu")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_lib
    oracle_lib.load()
    return oracle_lib
