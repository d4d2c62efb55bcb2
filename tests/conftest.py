import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def dj():
    """The product package (loads libdjets_b200.so; fails loudly when it has not been built)."""
    import djets_loader
    return djets_loader.load()


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure only)."""
    from oracle import djets_oracle
    return djets_oracle


@pytest.fixture(scope="session")
def ctx4(dj):
    """Four ranks driven by this process; on a 1-GPU box they share the GPU, each on its own stream,
    so grid exchanges (device copies ordered by events) run for real."""
    ctx = dj.init(world=4)
    yield ctx


@pytest.fixture(scope="session")
def ctx1(dj):
    ctx = dj.Context(world=1)
    yield ctx
