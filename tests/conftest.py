import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# tests written without GPU time to run them: parked until PDX_NEXT_ROUND=1 (tests/next_round/README.md)
collect_ignore_glob = [] if os.environ.get('PDX_NEXT_ROUND') == '1' else ['next_round/*']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda', 0)
