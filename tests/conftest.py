import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# tests/test_tiled_one_gpu.py runs several tiles on ONE GPU from threads of this process; their
# kernels wait for each other, so nothing may synchronise the whole context while they run: load
# every kernel up front (lazy loading synchronises on first use) and give the per-tile streams
# their own hardware queues.  Both must be set before CUDA initialises.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


from tealeaf_jl_b200.decks import CLASSIC_DECK, classic_settings  # noqa: E402,F401  (the deck lives in the package)


@pytest.fixture
def classic():
    return classic_settings
