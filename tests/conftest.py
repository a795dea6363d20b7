import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def fixture_data(oracle):
    return oracle.load_fixture()


@pytest.fixture(scope="session")
def ctx():
    import baynorm_b200 as bn
    c = bn.Context(0)          # raises if the CUDA library or the GPU is missing: no fallback
    yield c
    c.close()
