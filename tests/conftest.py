import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_numpy_half.npz'))


@pytest.fixture(scope='session')
def golden_edge():
    """degenerate lattices (oracle/make_golden.py edge): find_conn of the reference over ALL states"""
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_numpy_half_edge.npz'))
