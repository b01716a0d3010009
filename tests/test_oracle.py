"""Pins for the CPU oracle (the reference ships no tests or golden vectors: PARITY UNPINNED,
so the oracle is pinned by an independent NumPy twin, by invariants and by committed golden
values generated from it -- tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

import tealeaf_jl_b200 as tl
from oracle.np_twin import Twin
from oracle.oracle import OracleChunk

from conftest import classic_settings

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")


def run_oracle(s, stepwise=False, **kw):
    chunk, geom = tl.initialiseapp(s, backend=OracleChunk, **kw)
    recs, final = tl.diffuse(chunk, s, geom, stepwise=stepwise)
    return chunk, geom, recs, final


def interior(a, hd=2):
    return a[hd:-hd, hd:-hd]


def test_cg_iteration_counts_match_survey_appendix_c():
    # SURVEY.md Appendix C (throwaway NumPy probe of the intended algorithm)
    for n, want in ((64, [66, 64]), (128, [136, 131]), (256, [280, 272])):
        _, _, recs, _ = run_oracle(classic_settings(n))
        assert [r["iters"] for r in recs] == want


def test_literal_smvp_is_a_diagonal_operator():
    # Appendix A #1: the as-written smvp returns before the off-diagonal terms
    s = classic_settings(64)
    _, _, recs, _ = run_oracle(s, literal_smvp=True)
    assert recs[0]["iters"] < 12  # "converges" in a handful of iterations to a different answer


def test_cheby_and_ppcg_match_survey_appendix_c():
    _, _, recs, _ = run_oracle(classic_settings(128, steps=1, solver="cheby"))
    assert (recs[0]["cg_iters"], recs[0]["cheby_iters"], recs[0]["est_iters"]) == (30, 49, 35)
    _, _, recs, _ = run_oracle(classic_settings(128, steps=1, solver="ppcg"))
    assert (recs[0]["cg_iters"], recs[0]["cheby_iters"], recs[0]["inner_total"]) == (30, 6, 60)


@pytest.mark.parametrize("solver", ["cg", "cheby", "ppcg"])
def test_oracle_matches_numpy_twin(solver):
    n = 96
    s = classic_settings(n, ny=80, steps=1, solver=solver)
    chunk, geom = tl.initialiseapp(s, backend=OracleChunk)
    tl.haloupdate(chunk, s, 1, ["energy", "density"])
    tw = Twin(chunk.get_field("density"), chunk.get_field("energy"), s.halodepth, s.coefficient)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    if solver == "cg":
        got = chunk.cg_solve(s, rx, ry)
        want = tw.cg_solve(rx, ry, s.eps, s.maxiters)
    elif solver == "cheby":
        got = chunk.cheby_solve(s, rx, ry)
        want = tw.cheby_solve(rx, ry, s.eps, s.maxiters, s.presteps, s.epslim, s.errorswitch)
    else:
        got = chunk.ppcg_solve(s, rx, ry)
        want = tw.ppcg_solve(rx, ry, s.eps, s.maxiters, s.presteps, s.epslim, s.errorswitch, s.ppcginnersteps)
    for k in ("iters", "cg_iters", "cheby_iters"):
        if k in want:
            assert got[k] == want[k], k
    # element-wise quantities that do not depend on reduction order are bit-identical
    np.testing.assert_array_equal(chunk.get_field("kx"), tw.kx)
    np.testing.assert_array_equal(chunk.get_field("ky"), tw.ky)
    np.testing.assert_array_equal(chunk.get_field("u0"), tw.u0)
    # the rest agrees to rounding (serial vs pairwise sums feed alpha/beta)
    scale = np.abs(tw.u).max()
    assert np.abs(interior(chunk.get_field("u")) - interior(tw.u)).max() / scale < 1e-11
    assert np.abs(chunk.get_field("u") - tw.u).max() / scale < 1e-11  # halos follow the same rules


def test_first_matvec_is_bitwise_equal_to_twin():
    s = classic_settings(70, ny=50, steps=1)
    chunk, geom = tl.initialiseapp(s, backend=OracleChunk)
    tl.haloupdate(chunk, s, 1, ["energy", "density"])
    tw = Twin(chunk.get_field("density"), chunk.get_field("energy"), s.halodepth)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    rro_o = chunk.cg_init(s.coefficient, rx, ry)
    rro_t = tw.cg_init(rx, ry)
    for f in ("u", "w", "r", "p", "kx", "ky"):
        np.testing.assert_array_equal(chunk.get_field(f), getattr(tw, f), err_msg=f)
    assert abs(rro_o - rro_t) <= 1e-13 * abs(rro_t)


def test_operator_is_symmetric_and_conserves_energy():
    s = classic_settings(48, ny=40, steps=1)
    chunk, geom = tl.initialiseapp(s, backend=OracleChunk)
    tl.haloupdate(chunk, s, 1, ["energy", "density"])
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    chunk.cg_init(s.coefficient, rx, ry)
    rng = np.random.default_rng(0)
    hd = s.halodepth

    def apply(v):
        full = np.zeros((chunk.x, chunk.y), order="F")
        full[hd:-hd, hd:-hd] = v
        chunk.set_field("p", full)
        chunk.haloupdate(["p"], 1)
        chunk.cg_w()
        return interior(chunk.get_field("w"))

    a, b = rng.standard_normal((48, 40)), rng.standard_normal((48, 40))
    Aa, Ab = apply(a), apply(b)
    assert abs(np.sum(Aa * b) - np.sum(a * Ab)) < 1e-10 * abs(np.sum(Aa * b))   # <Aa,b> == <a,Ab>
    assert abs(np.sum(apply(np.ones((48, 40)))) - 48 * 40) < 1e-9               # rows of A - I sum to 0
    # a full solve conserves sum(u) (zero-flux boundaries)
    s2 = classic_settings(48, ny=40, steps=1)
    chunk2, geom2, recs, _ = run_oracle(s2)
    u, u0 = interior(chunk2.get_field("u")), interior(chunk2.get_field("u0"))
    assert abs(u.sum() - u0.sum()) < 1e-12 * abs(u0.sum())
    # CG's recurrence residual agrees with the true residual r = u0 - A u (residual!)
    r = interior(chunk2.get_field("r"))
    assert np.abs(r).max() < 1e-12


def test_solvers_agree_with_each_other():
    us = {}
    for solver in ("cg", "cheby", "ppcg"):
        chunk, _, _, final = run_oracle(classic_settings(64, steps=1, solver=solver))
        us[solver] = (interior(chunk.get_field("u")), final["temp"])
    scale = np.abs(us["cg"][0]).max()
    assert np.abs(us["cheby"][0] - us["cg"][0]).max() / scale < 1e-6
    assert np.abs(us["ppcg"][0] - us["cg"][0]).max() / scale < 1e-6
    assert abs(us["cheby"][1] / us["cg"][1] - 1) < 1e-8
    assert abs(us["ppcg"][1] / us["cg"][1] - 1) < 1e-8


@pytest.mark.parametrize("solver", ["cg", "cheby", "ppcg"])
def test_stepwise_python_driver_equals_c_driver(solver):
    # the host mirror of solve! (per-kernel calls) and the oracle's own C driver are the same algorithm
    a = run_oracle(classic_settings(64, steps=2, solver=solver), stepwise=False)
    b = run_oracle(classic_settings(64, steps=2, solver=solver), stepwise=True)
    assert [r["iters"] for r in a[2]] == [r["iters"] for r in b[2]]
    scale = np.abs(a[0].get_field("u")).max()
    assert np.abs(a[0].get_field("u") - b[0].get_field("u")).max() / scale < 1e-12


def test_recip_conductivity_and_rectangular_mesh():
    s = classic_settings(40, ny=72, steps=1, coefficient=tl.RECIP_CONDUCTIVITY)
    chunk, _, recs, final = run_oracle(s)
    assert recs[0]["iters"] > 5 and np.isfinite(final["temp"])
    with pytest.raises(ValueError):
        OracleChunk(8, 8).cg_init(3, 1.0, 1.0)   # CG.jl:48-50


def test_golden_values():
    with open(GOLDEN) as fh:
        golden = json.load(fh)
    for case in golden["cases"]:
        s = classic_settings(case["nx"], ny=case["ny"], steps=case["steps"], solver=case["solver"])
        _, _, recs, final = run_oracle(s)
        assert [r["iters"] for r in recs] == case["iters"], case
        for k in ("vol", "mass", "ie", "temp"):
            assert abs(final[k] - case["final"][k]) <= 1e-13 * abs(case["final"][k]), (case, k)


def test_jacobi_oracle_matches_numpy_twin_and_converges_towards_cg():
    """Jacobi: the C oracle and the NumPy twin take
    the same iterations element for element, the every-50th-iteration true residual replaces the
    error, and the (slow) iteration approaches the CG solution of the same system."""
    s = classic_settings(48, ny=40, steps=1, solver="jacobi", maxiters=73)
    chunk, geom = tl.initialiseapp(s, backend=OracleChunk)
    tl.haloupdate(chunk, s, 1, ["energy", "density"])
    tw = Twin(chunk.get_field("density"), chunk.get_field("energy"), s.halodepth, s.coefficient)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    got = chunk.jacobi_solve(s, rx, ry)
    want = tw.jacobi_solve(rx, ry, s.eps, s.maxiters)
    assert got["iters"] == want["iters"] == 73
    np.testing.assert_array_equal(chunk.get_field("kx"), tw.kx)
    np.testing.assert_array_equal(chunk.get_field("ky"), tw.ky)
    np.testing.assert_array_equal(chunk.get_field("u"), tw.u)       # element-wise: bit-identical
    np.testing.assert_array_equal(chunk.get_field("r"), tw.r)       # r = the previous iterate (Jacobi.jl:64)
    assert abs(got["error"] / want["error"] - 1) < 1e-12
    # at a multiple of 50 the error is the squared true residual and r holds the residual
    # (the classic deck is strongly diagonally dominant: the test sum(r.r) < eps fires at tt = 100)
    s2 = classic_settings(48, ny=40, steps=1, solver="jacobi", maxiters=230)
    c2, _ = tl.initialiseapp(s2, backend=OracleChunk)
    tl.haloupdate(c2, s2, 1, ["energy", "density"])
    g2 = c2.jacobi_solve(s2, rx, ry)
    assert g2["iters"] == 100 and tw.__class__(c2.get_field("density"), c2.get_field("energy"), s2.halodepth,
                                               s2.coefficient).jacobi_solve(rx, ry, s2.eps, 230)["iters"] == 100
    assert abs(g2["error"] - c2.norm2("r")) <= 1e-12 * g2["error"] and g2["error"] < s2.eps
    # stepwise host loop == the oracle's own driver
    s3 = classic_settings(48, ny=40, steps=1, solver="jacobi", maxiters=230)
    c3, _ = tl.initialiseapp(s3, backend=OracleChunk)
    tl.haloupdate(c3, s3, 1, ["energy", "density"])
    g3 = tl.get_solver("jacobi").solve_stepwise(c3, s3, rx, ry)
    assert g3["iters"] == 100 and g3["error"] == g2["error"]
    np.testing.assert_array_equal(c3.get_field("u"), c2.get_field("u"))
    # Jacobi and CG solve the same system: after many sweeps the iterate is close to CG's answer
    s4 = classic_settings(24, ny=24, steps=1, solver="jacobi", maxiters=600)
    c4, _ = tl.initialiseapp(s4, backend=OracleChunk)
    tl.haloupdate(c4, s4, 1, ["energy", "density"])
    c4.jacobi_solve(s4, s4.dtinit / s4.dx ** 2, s4.dtinit / s4.dy ** 2)
    s5 = classic_settings(24, ny=24, steps=1, solver="cg")
    c5, _ = tl.initialiseapp(s5, backend=OracleChunk)
    tl.haloupdate(c5, s5, 1, ["energy", "density"])
    c5.cg_solve(s5, s5.dtinit / s5.dx ** 2, s5.dtinit / s5.dy ** 2)
    a, b = interior(c4.get_field("u")), interior(c5.get_field("u"))
    assert np.abs(a - b).max() / np.abs(b).max() < 1e-12



def test_end_time_ends_the_timestep_loop():
    """`end_time` (settings.jl:58: parsed, never used by the reference) ends the loop once tt * dt reaches it, as
    upstream does; with the default 10.0 the decks' `end_step` decides."""
    import tealeaf_jl_b200 as tl
    from tealeaf_jl_b200.decks import CLASSIC_DECK
    from oracle.oracle import OracleChunk
    deck = CLASSIC_DECK.format(nx=24, ny=24, steps=10, solver="cg").replace("end_step=10", "end_step=10\nend_time=0.012")
    s = tl.parse_settings_text(deck)
    assert s.endtime == 0.012 and s.endstep == 10
    chunk, geom = tl.initialiseapp(s, backend=OracleChunk)
    recs, final = tl.diffuse(chunk, s, geom)
    assert len(recs) == 3                      # 3 * 0.004 >= 0.012
    s2 = tl.parse_settings_text(CLASSIC_DECK.format(nx=24, ny=24, steps=3, solver="cg"))
    chunk2, geom2 = tl.initialiseapp(s2, backend=OracleChunk)
    recs2, final2 = tl.diffuse(chunk2, s2, geom2)
    assert [r["iters"] for r in recs] == [r["iters"] for r in recs2] and final["temp"] == final2["temp"]


def test_tea_out_report(tmp_path):
    """`tea.out`-style report: every step, the four-component summaries, the QA verdict."""
    s = classic_settings(10, 10, steps=2, solver="cg")
    s.checkresult = True
    s.summaryfrequency = 1
    chunk, geom, recs, final = run_oracle(s)
    path = tmp_path / "tea.out"
    tl.write_tea_out(str(path), s, recs, final, wall_s=0.5)
    text = path.read_text()
    assert text.count(" Step ") == 2 and text.count("step:") == 2 and "final:" in text
    for r in recs:
        assert f" Iteration count {r['iters']:8d}" in text
    last = [ln for ln in text.splitlines() if ln.lstrip().startswith("final:")][0].split()
    assert float(last[1]) == pytest.approx(final["vol"], rel=1e-7) and float(last[5]) == pytest.approx(final["temp"], rel=1e-7)
    assert float(last[3]) == pytest.approx(final["mass"] / final["vol"], rel=1e-7)
    assert ("This test is considered PASSED" in text) == bool(final["passed"])
