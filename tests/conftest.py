"""pytest configuration: registers the `gpu` marker and exposes the package / oracle loaders."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def qmann():
    """The product package (directory q-mann_b200/, importable name qmann_b200)."""
    import __graft_entry__ as ge
    return ge.import_package()


@pytest.fixture(scope="session")
def synth(qmann):
    return qmann.synth


@pytest.fixture(scope="session")
def qmo():
    """ctypes binding of the CPU oracle (oracle/qmo.py) -- the checker, never the product."""
    import qmo as _qmo
    _qmo.lib()
    return _qmo
