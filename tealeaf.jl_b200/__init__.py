"""tealeaf.jl_b200 -- B200-native (sm_100a) drop-in for the implicit heat-conduction solve of
Laura7089/TeaLeaf.jl: the CG / Chebyshev / PPCG iterations on the 5-point stencil.

Contents (only what the hot path needs):
  csrc/        hand-written CUDA kernels + the C-ABI shared library (include/tealeaf_b200.h)
  lib.py       ctypes binding of that ABI (the Python twin of julia/TeaLeafB200.jl)
  device.py    DeviceChunk: the device-resident `Chunk` with the reference's kernel names
  settings.py  tea.in / tea.problems parser   (mirror of src/settings.jl)
  chunk.py     geometry + initial-state painter (mirror of src/chunk.jl)
  solvers.py   CG / Cheby / PPCG `solve!`      (mirror of src/solvers/*.jl)
  app.py       initialiseapp! / diffuse!       (mirror of src/TeaLeaf.jl)
  dist.py      2-D domain decomposition over torch.distributed (one process per GPU)
  run.py       CLI                             (mirror of run.jl)

The directory name contains a dot, so it is imported through the `tealeaf_jl_b200` shim at
the repository root.
"""
from .settings import Settings, State, parse_settings, parse_settings_text, checkingvalue, resettoexchange  # noqa: F401
from .settings import CONDUCTIVITY, RECIP_CONDUCTIVITY, EXCHANGE_FIELDS  # noqa: F401
from .chunk import HostGeometry, paint_states, FIELD_NAMES, FIELD_IDS  # noqa: F401
from .solvers import CG, Cheby, PPCG, get_solver, haloupdate  # noqa: F401
from .app import initialiseapp, diffuse, fieldsummary, upload_initial_state, write_tea_out  # noqa: F401


def DeviceChunk(*a, **kw):
    """Lazy constructor: loading the CUDA library is deferred until a device chunk is needed."""
    from .device import DeviceChunk as _D
    return _D(*a, **kw)
