import os
import sys

# The rank-thread tests drive several rank engines (4 streams each) inside this one process.  CUDA maps the streams of a
# process onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); streams that share a queue serialise, and a kernel
# queued behind another rank's spinning barrier kernel never starts.  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda_device():
    if not _cuda_available():
        pytest.skip("no CUDA device")
    return 0
