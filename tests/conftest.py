import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build the CUDA library and the C oracle once per session if they are missing."""
    from model_based_pde_control_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    from oracle import ks_c

    ks_c.build()


def load_golden(name):
    import numpy as np

    return np.load(os.path.join(GOLDEN, name + ".npz"))


STEP_CASES = [
    "kat1_default_1period", "kat2_default_10periods", "kat3_large_1period", "attractor_default_random",
    "attractor_default_zero_action", "attractor_default_saturated", "attractor_default_action1d",
    "truncation_edge", "attractor_large_random", "attractor_n128_random", "short_period_cfg10",
    "attractor_n96_random",
]
