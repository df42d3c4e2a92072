import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: about a minute each (full-size parity against the unmodified reference); still part of `-m gpu`")


@pytest.fixture(scope="session")
def built():
    """The native artefacts must exist; build them once if the tree is fresh."""
    lib = os.path.join(ROOT, "mapcaller_b200", "libmapcaller_b200.so")
    orc = os.path.join(ROOT, "oracle", "libmcoracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__ as g
        g.build()
    return True
