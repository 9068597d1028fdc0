import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built library (build products are not tracked): build it once with nvcc, exactly as
    __graft_entry__.build() does.  Building is not a fallback -- without nvcc the tests that need the library fail loudly."""
    import shutil
    import subprocess

    lib = os.path.join(ROOT, "forge-ec_b200", "lib", "libforge_ec_cuda.so")
    if not os.path.exists(lib) and shutil.which("nvcc") and shutil.which("make"):
        subprocess.run(["make", "-j", str(min(8, os.cpu_count() or 1)), "-C", os.path.join(ROOT, "forge-ec_b200", "csrc")], check=False,
                       stdout=subprocess.DEVNULL)


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
