import json
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
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    d = {k: z[k] for k in z.files}
    d['meta'] = json.loads(bytes(d['meta']).decode())
    return d


GOLDEN_CASES = ['small_onehot', 'small_onehot_odd', 'small_mbv', 'small_continues', 'small_gumbel_t', 'small_binary',
                'small_zeropad', 'full_b2_t128', 'full_b1_t207', 'full_b1_t9', 'full_b1_mbv', 'full_b1_e512']


@pytest.fixture(scope='session')
def golden():
    return load_golden
