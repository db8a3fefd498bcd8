import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_problem():
    """Config-1-like scene: shipped yml shape (T=20, D=7) on a 64^3 SDF."""
    from motion_planners_b200 import problems as P
    return P.single_arm_problem(K=10, T=20, sdf_n=64)


@pytest.fixture(scope="session")
def medium_problem():
    """Config-2-like scene at a size the CPU oracle finishes in seconds (K=32, T=100, 128^3)."""
    from motion_planners_b200 import problems as P
    return P.single_arm_problem(K=32, T=100, sdf_n=128)
