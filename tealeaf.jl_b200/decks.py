"""Synthetic input decks: the reference ships no tea.in (its .gitignore excludes
them), so the classic 5-state TeaLeaf benchmark geometry -- what upstream's tea_bm_* decks use --
is authored here in the reference's own grammar (src/settings.jl:90-173) and parsed by the same
parser mirror everything else goes through.  Deterministic, no RNG."""
from __future__ import annotations

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
    """The classic 5-state TeaLeaf benchmark deck at nx x ny; keyword
    overrides are `Settings` field names (maxiters, ppcginnersteps, halodepth, ...)."""
    from .settings import parse_settings_text
    s = parse_settings_text(CLASSIC_DECK.format(nx=nx, ny=ny or nx, steps=steps, solver=solver))
    for k, v in over.items():
        setattr(s, k, v)
    s.recompute_spacing()
    return s
