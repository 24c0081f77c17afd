import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    ge.build()
    return ge.load_package()


def _oracle_flavour():
    import refdrv
    return "strict" if refdrv.available("strict") else "port"


@pytest.fixture(scope="session")
def oracle_flavour(pkg):
    """The compiled reference when oracle/_ref/libref_oracle.so travelled with the repo, else the CPU restatement."""
    return _oracle_flavour()
