import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'fokl-gpy_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def cubic_table():
    return np.load(os.path.join(GOLD, 'phis_cubic_48.npy'))


@pytest.fixture(scope='session')
def phis_cubic(cubic_table):
    import spline_table
    return spline_table.to_phis(cubic_table)


@pytest.fixture(scope='session')
def bern_table():
    return np.load(os.path.join(GOLD, 'bernoulli_table.npy'))


@pytest.fixture(scope='session')
def phis_bern(bern_table):
    return tuple(list(bern_table[n, :n + 2]) for n in range(bern_table.shape[0]))


@pytest.fixture(scope='session')
def engine():
    from FoKL import FoKLRoutines
    return FoKLRoutines._engine()


def load_golden(name):
    return np.load(os.path.join(GOLD, name + '.npz'), allow_pickle=True)
