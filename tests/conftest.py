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


CLASSIC_DECK = """*tea
state 1 density=100.0 energy=0.0001
state 2 density=0.1 energy=25.0 geometry=rectangle xmin=0.0 xmax=1.0 ymin=1.0 ymax=2.0
state 3 density=0.1 energy=0.1 geometry=rectangle xmin=1.0 xmax=6.0 ymin=1.0 ymax=2.0
state 4 density=0.1 energy=0.1 geometry=rectangle xmin=5.0 xmax=6.0 ymin=1.0 ymax=8.0
state 5 density=0.1 energy=0.1 geometry=rectangle xmin=5.0 xmax=10.0 ymin=7.0 ymax=8.0
x_cells={nx}
y_cells={ny}
xmin=0.0
ymin=0.0
xmax=10.0
ymax=10.0
initial_timestep=0.004
end_step={steps}
max_iters=10000
use_{solver}
eps=1.0e-15
check_result=false
*endtea
"""


def classic_settings(nx, ny=None, steps=2, solver="cg", **over):
    """The classic 5-state TeaLeaf benchmark deck at nx x ny."""
    import tealeaf_jl_b200 as tl
    s = tl.parse_settings_text(CLASSIC_DECK.format(nx=nx, ny=ny or nx, steps=steps, solver=solver))
    for k, v in over.items():
        setattr(s, k, v)
    s.recompute_spacing()
    return s


@pytest.fixture
def classic():
    return classic_settings
