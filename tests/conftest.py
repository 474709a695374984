"""pytest configuration: the ``gpu`` marker and shared helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


def golden_callables(g):
    """(influence(dk), propagators(step)) callables from a golden fixture."""
    infl = g["influences"]

    def influence(dk):
        if dk < 0:
            return None
        return infl[dk]

    p1, p2 = g["prop_1"], g["prop_2"]

    def propagators(step):
        return p1, p2

    return influence, propagators
