import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure the product library and the oracle are compiled (cheap no-op when current)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle.oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def ref(built):
    from oracle.oracle import Ref

    if not Ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return Ref()


@pytest.fixture(scope="session")
def eng(built):
    from phylocaml_b200 import engine

    e = engine.Engine(0)
    yield e
    e.close()
