import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle_py as O
    O.lib()  # builds liboit_oracle.so if needed
    return O


@pytest.fixture(scope="session")
def oit_mod():
    import vk_order_independent_transparency_b200 as oit
    return oit
