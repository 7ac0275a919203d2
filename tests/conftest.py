import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ks():
    """The product package (requires libksched.so; built by __graft_entry__.build())."""
    import __graft_entry__ as g
    g.build()
    import ksched_pkg
    return ksched_pkg.load()


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle binding (test infrastructure)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import orc as _orc
    return _orc
