import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_sessionstart(session):
    """Native artefacts are git-ignored: build whatever is missing (nvcc cross-compiles without a GPU) so the
    suite does not depend on __graft_entry__.build() having run first."""
    import oat_b200
    import oracle

    if not os.path.exists(oat_b200.LIB_PATH):
        oat_b200.build()
    oracle.build()


@pytest.fixture(scope="session")
def ctx():
    import oat_b200

    c = oat_b200.Context(0)
    yield c
    c.close()
