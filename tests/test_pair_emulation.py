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
import emulate_pair_tiled  # noqa: E402
import emulate_ppcg_pair_tiled  # noqa: E402


@pytest.mark.parametrize("nx,ny,rows", emulate_pair.CASES)
def test_chebyshev_pair_window_logic(nx, ny, rows):
    emulate_pair.run(nx, ny, rows)


@pytest.mark.parametrize("nx,ny,rows", emulate_ppcg_pair.CASES)
def test_ppcg_pair_window_logic(nx, ny, rows):
    emulate_ppcg_pair.run(nx, ny, rows)


@pytest.mark.parametrize("NX,NY,px,py,rows", emulate_pair_tiled.CASES)
def test_tiled_pair_design_halo_depths_are_sufficient(NX, NY, px, py, rows):
    """the same window kernel on px x py tiles with u two cells deep (corners included),
    p / u0 one deep and kx / ky two deep in the tile-internal halos reproduces the single chunk."""
    assert emulate_pair_tiled.run(NX, NY, px, py, rows)


def test_tiled_pair_design_halo_depths_are_necessary():
    assert not emulate_pair_tiled.run(130, 12, 2, 2, 3, du=1)
    assert not emulate_pair_tiled.run(130, 12, 2, 2, 3, dp=0)
    assert not emulate_pair_tiled.run(130, 12, 2, 2, 3, dk=1)


@pytest.mark.parametrize("NX,NY,px,py,rows", emulate_ppcg_pair_tiled.CASES)
def test_tiled_ppcg_pair_design_halo_depths_are_sufficient(NX, NY, px, py, rows):
    """k_ppcg_pair_ring on tiles: sd two cells deep (corners included), r one deep, kx / ky two deep, nothing of u."""
    assert emulate_ppcg_pair_tiled.run(NX, NY, px, py, rows)


def test_tiled_ppcg_pair_design_halo_depths_are_necessary():
    assert not emulate_ppcg_pair_tiled.run(130, 12, 2, 2, 3, ds=1)
    assert not emulate_ppcg_pair_tiled.run(130, 12, 2, 2, 3, dr=0)
    assert not emulate_ppcg_pair_tiled.run(130, 12, 2, 2, 3, dk=1)


@pytest.mark.parametrize("first", [0, 1, 30])
def test_lazy_u_schedule_of_cg_kernel_a(first):
    """TL_U_LAZY: u advanced every second launch with both pending updates + the flush = an update per iteration, bit for bit,
    whatever the number of iterations and the parity of the phase's first iteration."""
    import emulate_lazy_u
    for n in range(0, 12):
        assert emulate_lazy_u.run(n, first), n
