import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from _libs import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from _libs import Ref, have_ref
    if not have_ref():
        pytest.skip("reference build (oracle/_ref) not available on this machine")
    return Ref()
