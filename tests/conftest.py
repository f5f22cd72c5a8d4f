import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


@pytest.fixture(scope='session')
def oracle():
    from oracle import ppsurf_oracle
    return ppsurf_oracle


@pytest.fixture(scope='session')
def weights(oracle):
    return oracle.make_state_dict(42)


@pytest.fixture(scope='session')
def weights_digest(oracle, weights):
    return oracle.state_dict_digest(weights)
