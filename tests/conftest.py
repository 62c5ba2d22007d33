import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_port():
    import oracle
    oracle.build(ref=True)
    return oracle.port()


@pytest.fixture(scope="session")
def oracle_ref():
    import oracle
    oracle.build(ref=True)
    if not oracle.Ref.available():
        pytest.skip("oracle/_ref not built (reference sources absent on this machine)")
    return oracle.ref()
