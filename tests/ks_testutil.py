"""Shared helpers for the test-suite (golden fixture access)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STEP_CASES = [
    "kat1_default_1period", "kat2_default_10periods", "kat3_large_1period", "attractor_default_random",
    "attractor_default_zero_action", "attractor_default_saturated", "attractor_default_action1d",
    "truncation_edge", "attractor_large_random", "attractor_n128_random", "short_period_cfg10",
    "attractor_n96_random",
]


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_l2(a, b):
    return np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)
