import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand from oracle/*.cpp."""
    from oracle import oracle as orc
    orc.build()
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def ctx():
    """One rl_ctx on cuda:0 for the whole GPU test session (fails loudly without the CUDA library)."""
    from rendering_learning_b200 import Context
    c = Context(0)
    yield c
    c.close()


GOLDEN = os.path.join(ROOT, "tests", "golden")
