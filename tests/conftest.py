import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fec():
    """the host package (hyphenated directory name -> importlib)"""
    return importlib.import_module("forge-ec_b200")


@pytest.fixture(scope="session")
def pm():
    """Python big-integer model of the path (oracle/pymodel.py) -- the checker, never the product"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pymodel

    return pymodel


@pytest.fixture(scope="session")
def engine(fec):
    """one context on cuda:0; a missing library or device is an error, not a skip"""
    eng = fec.Engine()
    yield eng
    eng.close()
