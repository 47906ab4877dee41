import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def ol():
    import oracle_lib
    oracle_lib.lib()      # builds oracle/liboracle.so on first use if it is missing
    return oracle_lib


@pytest.fixture(scope="session")
def rb(ol):
    return ol.rb
