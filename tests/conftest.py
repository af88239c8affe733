import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    # the oracle runs on torch-CPU: on many-core hosts the default thread count (all cores) is ~50x slower than 16
    # threads for these layer sizes (oversubscribed intra-op pool), so cap it for the whole session
    import torch
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))


@pytest.fixture(scope="session")
def repo_root():
    return ROOT
