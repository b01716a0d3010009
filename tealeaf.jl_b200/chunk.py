"""Host-side Chunk geometry and initial-state painter -- mirror of `src/chunk.jl`.

In the real integration Julia's `Chunk` / `setchunkstate!` do this work and the painted
`density` / `energy0` are handed to the library with `tl_set_field`.  The mirror is needed
here because Julia cannot run in the build/test environment.

Arrays are NumPy, Fortran order, shape (x, y) = (xcells + 2*hd, ycells + 2*hd), so that
`a[k, j]` (0-based) is the reference's `a[k+1, j+1]` and the memory layout equals Julia's.
A `tile` (x0, y0, nx, ny) paints the sub-rectangle of the global mesh a rank owns; the
geometry test is in physical coordinates (`src/chunk.jl:130-146`), so tiles are painted
independently.
"""
from __future__ import annotations

import numpy as np

from .settings import CIRCULAR, POINT, RECTANGULAR, Settings

FIELD_NAMES = ["density", "energy0", "energy", "u", "u0", "p", "r", "w", "kx", "ky", "sd"]
FIELD_IDS = {n: i for i, n in enumerate(FIELD_NAMES)}


class HostGeometry:
    """The coordinate part of `Chunk(settings)`, src/chunk.jl:68-89."""

    def __init__(self, settings: Settings, tile=None):
        hd = settings.halodepth
        if tile is None:
            tile = (0, 0, settings.xcells, settings.ycells)
        self.x0, self.y0, self.nx, self.ny = tile
        self.hd = hd
        self.x = self.nx + 2 * hd   # src/chunk.jl:69
        self.y = self.ny + 2 * hd   # src/chunk.jl:70
        # src/chunk.jl:76-77: vertexx[i] = xmin + dx*((1:x+1) - hd - 1), shifted by the tile origin
        ix = np.arange(self.x + 1, dtype=np.float64) + (self.x0 - hd)
        iy = np.arange(self.y + 1, dtype=np.float64) + (self.y0 - hd)
        self.vertexx = settings.xmin + settings.dx * ix
        self.vertexy = settings.ymin + settings.dy * iy
        # src/chunk.jl:43-44 (celly uses vertexy: Appendix A #4)
        self.cellx = 0.5 * (self.vertexx[:-1] + self.vertexx[1:])
        self.celly = 0.5 * (self.vertexy[:-1] + self.vertexy[1:])
        self.cell_volume = settings.dx * settings.dy  # src/chunk.jl:79


def paint_states(settings: Settings, geom: HostGeometry):
    """`setchunkstate!`, src/chunk.jl:122-151 (with [k,j] indexing, Appendix A #3).

    Returns (density, energy0, u) as Fortran-ordered (x, y) arrays."""
    x, y = geom.x, geom.y
    states = settings.states
    if not states:
        raise ValueError("no states in the deck")
    energy0 = np.full((x, y), states[0].energy, dtype=np.float64, order="F")   # :124
    density = np.full((x, y), states[0].density, dtype=np.float64, order="F")  # :125
    vx, vy = geom.vertexx, geom.vertexy
    for s in states:  # :130 -- all states, the first included, exactly as written
        if s.geometry == RECTANGULAR:
            mx = (vx[1:] >= s.xmin) & (vx[:-1] < s.xmax)
            my = (vy[1:] >= s.ymin) & (vy[:-1] < s.ymax)
            mask = mx[:, None] & my[None, :]
        elif s.geometry == CIRCULAR:
            mask = ((geom.cellx[:, None] - s.xmin) ** 2 + (geom.celly[None, :] - s.ymin) ** 2) <= s.radius ** 2
        elif s.geometry == POINT:
            mask = (vx[:-1] == s.xmin)[:, None] & (vy[:-1] == s.ymin)[None, :]
        else:
            raise ValueError(s.geometry)
        energy0[mask] = s.energy
        density[mask] = s.density
    u = np.zeros((x, y), dtype=np.float64, order="F")
    u[1:-1, 1:-1] = energy0[1:-1, 1:-1] * density[1:-1, 1:-1]  # :149-150, halo(ch, 1)
    return density, energy0, u


def reflect_halo_host(a: np.ndarray, hd: int, depth: int, phys=(True, True, True, True)) -> None:
    """Host version of updateface! (src/kernels.jl:191-210, Appendix A #2), used for tests."""
    x, y = a.shape
    for d in range(1, depth + 1):
        if phys[0]:
            a[hd - d, hd:y - hd] = a[hd + d - 1, hd:y - hd]
        if phys[1]:
            a[x - hd + d - 1, hd:y - hd] = a[x - hd - d, hd:y - hd]
    for d in range(1, depth + 1):
        if phys[3]:
            a[hd:x - hd, y - hd + d - 1] = a[hd:x - hd, y - hd - d]
        if phys[2]:
            a[hd:x - hd, hd - d] = a[hd:x - hd, hd + d - 1]
