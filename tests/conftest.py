import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    from sparta_b200 import build
    build.build()
    import sparta_b200
    return sparta_b200.load()


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_py import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.oracle_py import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return Reference()
