"""CPU check of the two-iterations-per-pass kernels' design (k_cheby_pair_ring, k_ppcg_pair_ring): NumPy
emulations of one warp task -- 64-column window over 60 owned columns, lane pairs, shuffles, the
row carry, the reflective clamps, garbage in every cell the kernel claims never to use -- must
reproduce two applications of the one-iteration formula bit for bit.  (The CUDA kernels themselves
are compared with the one-iteration kernels in tests/test_gpu_parity.py, -m gpu.)"""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emulation"))
import emulate_pair  # noqa: E402
import emulate_ppcg_pair  # noqa: E402


@pytest.mark.parametrize("nx,ny,rows", emulate_pair.CASES)
def test_chebyshev_pair_window_logic(nx, ny, rows):
    emulate_pair.run(nx, ny, rows)


@pytest.mark.parametrize("nx,ny,rows", emulate_ppcg_pair.CASES)
def test_ppcg_pair_window_logic(nx, ny, rows):
    emulate_ppcg_pair.run(nx, ny, rows)
