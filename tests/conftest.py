import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle import pyoracle
    if not pyoracle.have_reference():
        if os.path.isdir("/root/reference/degensac"):
            pyoracle.build(with_ref=True)
        else:
            pytest.skip("oracle/_ref/libmods_ref.so not built (needs /root/reference)")
    return pyoracle.Reference()


@pytest.fixture(scope="session")
def ctx():
    import mods_b200 as mb
    c = mb.Context(0)  # raises without a GPU: there is no CPU fallback
    yield c
    c.close()
