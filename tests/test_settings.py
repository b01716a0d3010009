"""Deck grammar and the initial-state painter (src/chunk.jl:122-151)."""
import numpy as np

import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.chunk import HostGeometry, paint_states, reflect_halo_host
from conftest import CLASSIC_DECK, classic_settings


def test_defaults_match_reference():
    s = tl.Settings()   # src/settings.jl:41-77
    assert (s.presteps, s.maxiters, s.coefficient, s.ppcginnersteps, s.summaryfrequency, s.halodepth) == \
        (30, 10_000, 1, 10, 10, 2)
    assert (s.eps, s.dtinit, s.epslim, s.errorswitch, s.checkresult) == (1e-15, 0.1, 1e-5, False, True)
    assert (s.xcells, s.ycells, s.xmin, s.xmax, s.solver) == (10, 10, 0.0, 100.0, "cg")
    assert set(s.toexchange) == {"density", "p", "energy0", "energy", "u", "sd"}


def test_classic_deck_parses():
    s = tl.parse_settings_text(CLASSIC_DECK.format(nx=40, ny=20, steps=3, solver="ppcg"))
    assert (s.xcells, s.ycells, s.endstep, s.solver, s.dtinit) == (40, 20, 3, "ppcg", 0.004)
    assert (s.dx, s.dy) == (0.25, 0.5)
    assert len(s.states) == 5 and s.states[0].density == 100.0 and s.states[1].energy == 25.0
    # the nudge uses the final dx, dy (Appendix A #6), state 1 has no geometry
    assert s.states[1].xmin == 0.0 + 0.25 / 100 and s.states[1].ymax == 2.0 - 0.5 / 100
    assert s.states[0].xmax == 0.0


def test_grammar_quirks():
    text = """*tea
state 1 density=1.0 energy=2.0
state 2 density=3.0 energy=4.0 geometry=circular xmin=5.0 ymin=5.0 radius=2.5
state 3 density=5.0 energy=6.0 geometry=point xmin=2.0 ymin=2.0
 x_cells = 20
y_cells=20
tl_max_iters=77
tl_use_chebyshev
tl_ppcg_inner_steps=4
tl_ppcg_halo_depth=2
test_problem 5
unknown_key=3
error_switch=true
check_result=false
coefficient=2
xmax=10.0
ymax=10.0
end_step=1
*endtea
"""
    s = tl.parse_settings_text(text)
    assert s.xcells == 20 and s.maxiters == 77 and s.solver == "cheby" and s.ppcginnersteps == 4
    assert s.errorswitch is True and s.checkresult is False and s.coefficient == 2
    assert s.ppcghalodepth == 2 and tl.Settings().ppcghalodepth == 0   # extension: depth-k tile exchange
    assert s.states[1].geometry == "circular" and s.states[1].radius == 2.5
    assert s.states[2].geometry == "point"
    s2 = tl.parse_settings_text("use_cg\nuse_jacobi\nuse_nonsense\n")
    assert s2.solver == "jacobi"


def test_tea_problems_lookup(tmp_path):
    p = tmp_path / "tea.problems"
    p.write_text("* comment\n\n10 10 2 1.5\n64 64 2 121.62583844121953\n")
    s = classic_settings(64, steps=2)
    assert tl.checkingvalue(s, str(p)) == 121.62583844121953
    s.endstep = 3
    assert tl.checkingvalue(s, str(p)) == 1.0            # not listed -> 1.0 (settings.jl:193-194)
    assert tl.checkingvalue(s, str(tmp_path / "missing")) == 1.0


def test_painter_rectangles_and_tiles():
    s = classic_settings(40, ny=30)
    g = HostGeometry(s)
    d, e, u = paint_states(s, g)
    assert d.shape == (44, 34) and d.flags.f_contiguous
    hd = 2
    # state 2: x in [0,1], y in [1,2] -> cells 0..3 in x (dx = .25), 3..5 in y (dy = 1/3)
    assert e[hd + 0, hd + 3] == 25.0 and d[hd + 0, hd + 3] == 0.1
    assert e[hd + 20, hd + 20] == 0.1 and d[hd + 20, hd + 20] == 0.1     # state 4 column
    assert d[hd + 10, hd + 20] == 100.0 and e[hd + 10, hd + 20] == 1e-4  # background
    np.testing.assert_array_equal(u[1:-1, 1:-1], (e * d)[1:-1, 1:-1])
    assert (u[0, :] == 0).all() and (u[:, -1] == 0).all()
    # a tile paints exactly the matching window of the global mesh
    tile = HostGeometry(s, tile=(8, 5, 16, 10))
    dt, et, _ = paint_states(s, tile)
    np.testing.assert_array_equal(dt, d[8:8 + 16 + 2 * hd, 5:5 + 10 + 2 * hd])
    np.testing.assert_array_equal(et, e[8:8 + 16 + 2 * hd, 5:5 + 10 + 2 * hd])


def test_painter_circle_and_point():
    text = ("state 1 density=1.0 energy=1.0\n"
            "state 2 density=2.0 energy=3.0 geometry=circular xmin=5.0 ymin=5.0 radius=2.0\n"
            "state 3 density=7.0 energy=9.0 geometry=point xmin=0.0 ymin=0.0\n"
            "x_cells=20\ny_cells=20\nxmax=10.0\nymax=10.0\nend_step=1\n")
    s = tl.parse_settings_text(text)
    g = HostGeometry(s)
    d, e, _ = paint_states(s, g)
    cx, cy = g.cellx[:, None], g.celly[None, :]
    inside = (cx - s.states[1].xmin) ** 2 + (cy - s.states[1].ymin) ** 2 <= 4.0
    assert (d[inside] == 2.0).all() and inside.sum() > 30
    # the point state is nudged by dx/100 like every geometry key (settings.jl:158), so it hits no vertex
    assert (d == 7.0).sum() == 0


def test_reflect_halo_host():
    a = np.arange(7 * 6, dtype=float).reshape(7, 6, order="F")
    b = a.copy(order="F")
    reflect_halo_host(b, 2, 2)
    assert (b[1, 2:-2] == a[2, 2:-2]).all() and (b[0, 2:-2] == a[3, 2:-2]).all()
    assert (b[-2, 2:-2] == a[-3, 2:-2]).all() and (b[2:-2, 0] == a[2:-2, 3]).all()
    assert b[0, 0] == a[0, 0]   # corners untouched, like the reference


def test_debugrecord_dump_follows_the_reference_format(tmp_path):
    """`--debug-out` (run.jl:20-22, TeaLeaf.jl:63-68, :90-107): one block per Chunk attribute in
    declaration order, matrices one column (fixed jj) per line, Julia float formatting; the file is
    restarted by diffuse! and appended to at every timestep.  Runs on the CPU oracle backend here;
    the device backend goes through the same code with tl_get_field."""
    import tealeaf_jl_b200 as tl
    from tealeaf_jl_b200.app import julia_float
    from oracle.oracle import OracleChunk
    from conftest import classic_settings
    assert [julia_float(v) for v in (100.0, 1e-4, 1e-5, 0.1, 1e6, -2.5e10, 0.0)] == \
        ["100.0", "0.0001", "1.0e-5", "0.1", "1.0e6", "-2.5e10", "0.0"]
    out = tmp_path / "debug.txt"
    out.write_text("stale")
    s = classic_settings(6, ny=5, steps=2, solver="cg")
    s.debugfile = str(out)
    chunk, geom = tl.initialiseapp(s, backend=OracleChunk)
    tl.diffuse(chunk, s, geom)
    text = out.read_text(encoding="utf-8")
    assert "stale" not in text
    blocks = text.split("\n\n\n\n")
    assert len([b for b in blocks if b.strip()]) == 2            # one record per timestep
    names = [blk.split("\n")[0] for blk in blocks[0].split("\n\n") if blk.strip()]
    assert names == ["density0", "density", "energy0", "energy", "u", "u0", "p", "r", "mi", "w", "kx", "ky", "sd",
                     "vertexx", "vertexy", "cellx", "celly", "volume", "θ", "eigmin", "eigmax", "cgα", "cgβ",
                     "chebyα", "chebyβ"]
    dens = [blk for blk in blocks[0].split("\n\n") if blk.startswith("density\n")][0].split("\n")[1:]
    assert len(dens) == 5 + 4 and all(len(row.split(" ")) == 6 + 4 for row in dens)
    assert {v for row in dens for v in row.split(" ")} == {"100.0", "0.1"}



def test_literal_nudge_switch_reproduces_the_unpatched_reference():
    """`readstate` runs inside the line loop, before dx, dy are recomputed (src/settings.jl:98-100 vs :132-133), so the
    literal reference always nudges the state rectangles by the DEFAULT dx/100 = 10/100 = 0.1.  The
    mirror's default is the intended two-pass nudge; literal_nudge=True reproduces the unpatched behaviour."""
    import tealeaf_jl_b200 as tl
    from tealeaf_jl_b200.decks import CLASSIC_DECK
    text = CLASSIC_DECK.format(nx=200, ny=100, steps=1, solver="cg")
    s = tl.parse_settings_text(text)
    lit = tl.parse_settings_text(text, literal_nudge=True)
    assert abs(s.dx - 0.05) < 1e-15 and abs(s.dy - 0.1) < 1e-15
    st, sl = s.states[1], lit.states[1]             # state 2: xmin=0.0 xmax=1.0 ymin=1.0 ymax=2.0
    assert abs(st.xmin - 0.0005) < 1e-15 and abs(st.xmax - 0.9995) < 1e-12 and abs(st.ymin - 1.001) < 1e-12
    assert abs(sl.xmin - 0.1) < 1e-15 and abs(sl.xmax - 0.9) < 1e-15 and abs(sl.ymin - 1.1) < 1e-15 and abs(sl.ymax - 1.9) < 1e-15
    lit.xcells = 400                                 # -x / -y overrides recompute the spacing, not the literal nudge
    lit.recompute_spacing()
    assert abs(lit.states[1].xmin - 0.1) < 1e-15
    # the painted regions differ: what the Julia-hosted path paints without the two-pass fix
    from tealeaf_jl_b200.chunk import HostGeometry, paint_states
    lit = tl.parse_settings_text(text, literal_nudge=True)
    d1, _, _ = paint_states(s, HostGeometry(s))
    d2, _, _ = paint_states(lit, HostGeometry(lit))
    assert (d1 != d2).sum() > 0
