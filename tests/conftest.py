import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as entry
    return entry.load_package()


@pytest.fixture(scope="session")
def engine(pkg):
    eng = pkg.Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def oracle_lib():
    """The oracle's C twin over GMP (built by __graft_entry__.build())."""
    import ctypes
    import __graft_entry__ as entry
    if not os.path.exists(entry.ORACLE_LIB):
        entry.build()
    lib = ctypes.CDLL(entry.ORACLE_LIB)
    lib.oracle_gmp_version.restype = ctypes.c_char_p
    return lib


@pytest.fixture(scope="session")
def keyset():
    from tests.golden import fixtures
    return fixtures.load_keyset()
