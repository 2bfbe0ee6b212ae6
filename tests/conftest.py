import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box with -m gpu)')


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def ld_join(z, key):
    """Rebuild a float128 vector stored as a float64 (hi, lo) pair."""
    import numpy as np
    return z[key + '_hi'].astype(np.longdouble) + z[key + '_lo'].astype(np.longdouble)
