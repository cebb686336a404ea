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
    import oracle_binding

    oracle_binding.lib()
    return oracle_binding


@pytest.fixture(scope="session")
def handle_factory():
    """Creates ls2d handles on cuda:0; fails loudly (no skip, no fallback) when CUDA is unavailable."""
    from srrg2_laser_slam_2d_b200 import Handle

    made = []

    def make(params=None):
        h = Handle(0, params)
        made.append(h)
        return h

    yield make
    for h in made:
        h.close()
