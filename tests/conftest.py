import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=True)


def tape_from_seed(seed, n):
    """The uniforms the reference consumed: numpy's MT19937 RandomState(seed).random_sample()."""
    return np.random.RandomState(int(seed)).random_sample(int(n))


@pytest.fixture(scope="session")
def oracle():
    from oracle import sda_oracle
    sda_oracle.build()
    return sda_oracle


def has_cuda():
    try:
        from sdaopt_b200 import _lib
        return _lib.lib().sda_device_count() > 0
    except Exception:
        return False
