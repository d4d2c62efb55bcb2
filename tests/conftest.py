import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` under gpurun)")
    # a fresh checkout has no libdjets_b200.so (built artefacts are git-ignored): build it once, with the
    # same recipe as __graft_entry__.build(); a failed build surfaces as the package's loud ImportError
    lib = os.path.join(ROOT, "distributedjets.jl_b200", "libdjets_b200.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "distributedjets.jl_b200", "csrc")], check=False)


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
