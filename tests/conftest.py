import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, ROOT)

MEDIA = os.path.join(ROOT, "tests", "data", "media")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """build liboracle.so, libgpurt.so (cross-compiles without a GPU) and the emu harness"""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def gpurt(built):
    import gpurt
    return gpurt


@pytest.fixture(scope="session")
def orc(built):
    import orc
    return orc


@pytest.fixture(scope="session")
def ctx(gpurt):
    c = gpurt.Context(0)
    yield c
    c.close()
