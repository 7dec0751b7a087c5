import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def load_isprs(name):
    import numpy as np
    d = np.load(os.path.join(GOLDEN, 'isprs_%s.npz' % name))
    return d['x'] / 100.0, d['y'] / 100.0, d['z'] / 100.0, d['g']


@pytest.fixture(scope='session')
def expected():
    import json
    with open(os.path.join(GOLDEN, 'isprs_expected.json')) as f:
        return json.load(f)
