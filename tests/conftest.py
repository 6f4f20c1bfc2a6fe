import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Built on demand from oracle/."""
    from oracle import binding

    binding.build()
    binding.load()
    return binding


@pytest.fixture(scope="session")
def vrt():
    """The product package; on a GPU box the CUDA library must load (no fallback)."""
    import voxel_rt2_b200

    return voxel_rt2_b200
