"""Parity of the CUDA path (through the C-ABI) against the CPU oracle.  Needs a B200: -m gpu.

Bars (BASELINE.json north_star): identical iteration count (+-1 where reduction order
differs), per-step field summaries within 1e-10 relative, final u within 1e-9 max relative
error.  Element-wise kernels are held to bit-exactness (same expression order, no FMA
contraction on either side)."""
import json
import os

import numpy as np
import pytest

import tealeaf_jl_b200 as tl
from conftest import classic_settings

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")
U_TOL = 1e-9        # final u, max relative error (north_star)
SUM_TOL = 1e-11     # dot products: tree (GPU) vs serial (oracle) summation order
SUMMARY_TOL = 1e-10  # vol / mass / ie / temp, relative (north_star)


def _oracle():
    from oracle.oracle import OracleChunk
    return OracleChunk


def _device():
    from tealeaf_jl_b200.device import DeviceChunk
    return DeviceChunk


def run(backend, s, stepwise=False):
    chunk, geom = tl.initialiseapp(s, backend=backend)
    summaries = []
    recs, final = tl.diffuse(chunk, s, geom, stepwise=stepwise,
                             on_step=lambda rec: summaries.append(chunk.fieldsummary(geom.cell_volume)))
    return chunk, recs, final, summaries


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def assert_parity(dev, ora, iter_slack=1, u_tol=U_TOL, fields=("u", "energy"), aux=()):
    dchunk, drecs, dfinal, dsum = dev
    ochunk, orecs, ofinal, osum = ora
    for dr, orr in zip(drecs, orecs):
        assert abs(dr["iters"] - orr["iters"]) <= iter_slack, (dr, orr)
    for ds, os_ in zip(dsum, osum):
        for a, b in zip(ds, os_):
            assert abs(a - b) <= SUMMARY_TOL * abs(b), (ds, os_)
    for f in fields:
        assert rel(dchunk.get_field(f), ochunk.get_field(f)) < u_tol, f
    # work vectors (p, r, w, sd) decay towards 0 at convergence, so they are compared on the scale of u
    scale = np.abs(ochunk.get_field("u")).max()
    for f in aux:
        assert np.abs(dchunk.get_field(f) - ochunk.get_field(f)).max() / scale < u_tol, f


# ---------------------------------------------------------------------------------------------
# per-function kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nx,ny,hd,coef", [(64, 64, 2, 1), (70, 50, 2, 1), (129, 65, 2, 2), (10, 10, 2, 1),
                                           (96, 40, 3, 2), (2, 300, 2, 1), (40, 3, 2, 1)])
# halo_depth = 1 is not a usable setting of the reference: CG.init! (CG.jl:57-63) reads w[kk-1] one cell
# outside the ring it initialised, giving Inf coefficients; the default is 2 (settings.jl:47).
def test_kernels_bitwise(nx, ny, hd, coef):
    s = classic_settings(nx, ny=ny, steps=1, halodepth=hd, coefficient=coef)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    chunks = []
    for backend in (_device(), _oracle()):
        c, geom = tl.initialiseapp(s, backend=backend)
        tl.haloupdate(c, s, 1, ["energy", "density"])
        chunks.append(c)
    d, o = chunks
    for f in ("density", "energy0", "energy"):
        np.testing.assert_array_equal(d.get_field(f), o.get_field(f), err_msg=f)
    rro_d, rro_o = d.cg_init(coef, rx, ry), o.cg_init(coef, rx, ry)
    for f in ("u", "w", "r", "p", "kx", "ky"):
        np.testing.assert_array_equal(d.get_field(f), o.get_field(f), err_msg=f"cg_init {f}")
    assert abs(rro_d - rro_o) <= SUM_TOL * abs(rro_o)
    for c in (d, o):
        c.haloupdate(["u", "p"], 1)
        c.copyu()
    np.testing.assert_array_equal(d.get_field("p"), o.get_field("p"))
    np.testing.assert_array_equal(d.get_field("u0"), o.get_field("u0"))
    pw_d, pw_o = d.cg_w(), o.cg_w()
    np.testing.assert_array_equal(d.get_field("w"), o.get_field("w"), err_msg="cg_w")
    assert abs(pw_d - pw_o) <= SUM_TOL * abs(pw_o)
    alpha = rro_o / pw_o
    rrn_d, rrn_o = d.cg_ur(alpha), o.cg_ur(alpha)
    np.testing.assert_array_equal(d.get_field("u"), o.get_field("u"), err_msg="cg_ur u")
    np.testing.assert_array_equal(d.get_field("r"), o.get_field("r"), err_msg="cg_ur r")
    assert abs(rrn_d - rrn_o) <= SUM_TOL * abs(rrn_o)
    beta = rrn_o / rro_o
    for c in (d, o):
        c.cg_p(beta)
        c.haloupdate(["u", "p"], 1)
    np.testing.assert_array_equal(d.get_field("p"), o.get_field("p"), err_msg="cg_p")
    # Chebyshev / PPCG building blocks with arbitrary coefficients
    theta = 3.0
    bb_d, bb_o = d.cheby_init(theta), o.cheby_init(theta)
    assert abs(bb_d - bb_o) <= SUM_TOL * abs(bb_o)
    e_d, e_o = d.cheby_iterate(0.4, 0.03, True, 0.0), o.cheby_iterate(0.4, 0.03, True, 0.0)
    assert abs(e_d - e_o) <= SUM_TOL * abs(e_o)
    for f in ("u", "w", "r", "p"):
        np.testing.assert_array_equal(d.get_field(f), o.get_field(f), err_msg=f"cheby {f}")
    for c in (d, o):
        c.ppcg_init_sd(theta)
        c.ppcg_inner([0.3, 0.2, 0.1], [0.05, 0.04, 0.03], 3)
    for f in ("u", "r", "sd"):
        np.testing.assert_array_equal(d.get_field(f), o.get_field(f), err_msg=f"ppcg {f}")
    assert abs(d.norm2("r") - o.norm2("r")) <= SUM_TOL * abs(o.norm2("r"))
    for c in (d, o):
        c.solvefinished(True)
    for f in ("r", "energy"):
        np.testing.assert_array_equal(d.get_field(f), o.get_field(f), err_msg=f"solvefinished {f}")
    sd, so = d.fieldsummary(0.37), o.fieldsummary(0.37)
    for a, b in zip(sd, so):
        assert abs(a - b) <= SUM_TOL * abs(b)


def test_halo_update_depths_and_roundtrip():
    D = _device()
    rng = np.random.default_rng(1)
    for hd, depth in ((2, 1), (2, 2), (3, 3), (1, 1), (4, 2)):
        nx, ny = 37, 21
        c = D(nx, ny, hd, 100)
        a = np.asfortranarray(rng.standard_normal((nx + 2 * hd, ny + 2 * hd)))
        c.set_field("sd", a)
        np.testing.assert_array_equal(c.get_field("sd"), a)   # H2D/D2H round trip, halos included
        c.haloupdate(["sd"], depth)
        want = a.copy(order="F")
        from tealeaf_jl_b200.chunk import reflect_halo_host
        reflect_halo_host(want, hd, depth)
        np.testing.assert_array_equal(c.get_field("sd"), want)
        c.close()


def test_errors_are_reported_not_thrown():
    from tealeaf_jl_b200.lib import TeaLeafError
    c = _device()(16, 16, 2, 50)
    with pytest.raises(TeaLeafError) as e:
        c.cg_init(3, 1.0, 1.0)            # CG.jl:48-50
    assert "Coefficient 3 is not valid" in str(e.value)
    with pytest.raises(TeaLeafError):
        c.haloupdate(["u"], 5)            # deeper than the halo (A#2: would index out of bounds)
    c.close()


MIXED_DECK = """*tea
state 1 density=100.0 energy=0.0001
state 2 density=0.1 energy=25.0 geometry=rectangle xmin=0.0 xmax=1.0 ymin=1.0 ymax=2.0
state 3 density=5.0 energy=2.5 geometry=circular xmin=6.0 ymin=4.0 radius=2.25
state 4 density=0.3 energy=7.0 geometry=point xmin=2.5 ymin=7.5
state 5 density=0.1 energy=0.1 geometry=rectangle xmin=5.0 xmax=10.0 ymin=7.0 ymax=8.0
x_cells={nx}
y_cells={ny}
xmin=0.0
ymin=0.0
xmax=10.0
ymax=10.0
initial_timestep=0.004
end_step=1
use_cg
*endtea
"""


@pytest.mark.parametrize("nx,ny,tile", [(40, 40, None), (97, 61, None), (80, 64, (40, 32, 40, 32)), (80, 64, (0, 32, 27, 32))])
def test_device_painter_is_bit_identical_to_the_host_painter(nx, ny, tile):
    """tl_paint_states (setchunkstate!, src/chunk.jl:122-151 on the device) vs the host mirror:
    rectangle, circular and point states, whole mesh and tiles with an offset."""
    from tealeaf_jl_b200.chunk import HostGeometry, paint_states
    for deck in (None, MIXED_DECK):
        s = classic_settings(nx, ny=ny, steps=1) if deck is None else tl.parse_settings_text(deck.format(nx=nx, ny=ny))
        geom = HostGeometry(s, tile=tile)
        d = _device()(geom.nx, geom.ny, s.halodepth, 100)
        d.paint_states(s, geom)
        density, energy0, u = paint_states(s, geom)
        np.testing.assert_array_equal(d.get_field("density"), density)
        np.testing.assert_array_equal(d.get_field("energy0"), energy0)
        np.testing.assert_array_equal(d.get_field("u"), u)
        d.close()
    # and the whole initialiseapp! agrees whichever side paints
    s = classic_settings(64, ny=48, steps=1)
    a, _ = tl.initialiseapp(s, backend=_device())
    b = _device()(64, 48, s.halodepth, s.maxiters)
    tl.upload_initial_state(b, s, host_paint=True)
    for f in ("density", "energy0", "energy", "u"):
        np.testing.assert_array_equal(a.get_field(f), b.get_field(f), err_msg=f)
    a.close(); b.close()


# ---------------------------------------------------------------------------------------------
# whole solves
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver", ["cg", "cheby", "ppcg"])
def test_stepwise_device_matches_oracle(solver):
    s = lambda: classic_settings(64, ny=48, steps=1, solver=solver)
    dev = run(_device(), s(), stepwise=True)
    ora = run(_oracle(), s(), stepwise=True)
    assert_parity(dev, ora, iter_slack=1, aux=("p", "w", "r"))


@pytest.mark.parametrize("nx,ny", [(64, 64), (128, 128), (200, 120), (10, 10), (65, 33), (257, 19), (1, 40), (3, 3)])
def test_fused_cg_matches_oracle(nx, ny):
    s = lambda: classic_settings(nx, ny=ny, steps=2, solver="cg")
    dev = run(_device(), s())
    ora = run(_oracle(), s())
    assert_parity(dev, ora, fields=("u", "energy", "u0", "kx", "ky"), aux=("p", "w"))
    # the true residual (residual!, run by solvefinished!) is tiny on both
    assert np.abs(dev[0].get_field("r")).max() < 1e-10


@pytest.mark.parametrize("nx,ny", [(128, 128), (96, 160), (65, 70)])
def test_fused_cheby_matches_oracle(nx, ny):
    s = lambda: classic_settings(nx, ny=ny, steps=2, solver="cheby")
    dev = run(_device(), s())
    ora = run(_oracle(), s())
    for dr, orr in zip(dev[1], ora[1]):
        assert dr["cg_iters"] == orr["cg_iters"]
        assert dr["est_iters"] == orr["est_iters"]
        assert abs(dr["eigmin"] / orr["eigmin"] - 1) < 1e-10 and abs(dr["eigmax"] / orr["eigmax"] - 1) < 1e-10
    assert_parity(dev, ora, iter_slack=0, aux=("p", "w", "r"))


@pytest.mark.parametrize("nx,ny,inner", [(128, 128, 10), (96, 160, 4), (65, 70, 7)])
def test_fused_ppcg_matches_oracle(nx, ny, inner):
    s = lambda: classic_settings(nx, ny=ny, steps=2, solver="ppcg", ppcginnersteps=inner)
    dev = run(_device(), s())
    ora = run(_oracle(), s())
    for dr, orr in zip(dev[1], ora[1]):
        assert dr["cg_iters"] == orr["cg_iters"]
        assert dr["inner_total"] == orr["inner_total"]
    assert_parity(dev, ora, iter_slack=0, aux=("p", "sd", "w"))


@pytest.mark.parametrize("nx,ny,cap", [(64, 48, 73), (131, 77, 100), (200, 120, 1000), (3, 3, 60), (1024, 640, 400)])
def test_jacobi_matches_oracle(nx, ny, cap):
    """Jacobi: per-function kernels bit-exact, fused solve element-wise
    bit-exact too (the sweep has no data-dependent control flow before the iteration cap)."""
    s = lambda: classic_settings(nx, ny=ny, steps=1, solver="jacobi", maxiters=cap)
    # per-function path, kernel by kernel
    d, _ = tl.initialiseapp(s(), backend=_device())
    o, _ = tl.initialiseapp(s(), backend=_oracle())
    st = s()
    rx, ry = st.dtinit / st.dx ** 2, st.dtinit / st.dy ** 2
    for c in (d, o):
        tl.haloupdate(c, st, 1, ["energy", "density"])
        c.jacobi_init(st.coefficient, rx, ry)
    for f in ("u", "u0", "kx", "ky"):
        np.testing.assert_array_equal(d.get_field(f), o.get_field(f), err_msg=f)
    for _ in range(3):
        ed, eo = d.jacobi_iterate(), o.jacobi_iterate()
        assert abs(ed - eo) <= SUM_TOL * abs(eo)
        d.haloupdate(["u"], 1); o.haloupdate(["u"], 1)
        np.testing.assert_array_equal(d.get_field("u"), o.get_field("u"))
        np.testing.assert_array_equal(d.get_field("r"), o.get_field("r"))
    d.close(); o.close()
    # whole solves: fused device path vs oracle, and the stepwise host loop on the device
    dev, ora = run(_device(), s()), run(_oracle(), s())
    assert dev[1][0]["iters"] == ora[1][0]["iters"] <= cap     # converges on a 50th-iteration residual or hits the cap
    assert abs(dev[1][0]["error"] - ora[1][0]["error"]) <= 1e-9 * abs(ora[1][0]["error"])
    np.testing.assert_array_equal(dev[0].get_field("u"), ora[0].get_field("u"))
    np.testing.assert_array_equal(dev[0].get_field("energy"), ora[0].get_field("energy"))
    assert_parity(dev, ora, iter_slack=0)
    step = run(_device(), s(), stepwise=True)
    np.testing.assert_array_equal(step[0].get_field("u"), ora[0].get_field("u"))


def test_jacobi_post_solve_r_and_coefficient_check():
    for cap in (49, 50):     # r = previous iterate / r = the 50th-iteration residual
        s = lambda: classic_settings(40, ny=56, steps=1, solver="jacobi", maxiters=cap, checkresult=False)
        d, _ = tl.initialiseapp(s(), backend=_device())
        o, _ = tl.initialiseapp(s(), backend=_oracle())
        st = s()
        rx, ry = st.dtinit / st.dx ** 2, st.dtinit / st.dy ** 2
        for c in (d, o):
            tl.haloupdate(c, st, 1, ["energy", "density"])
        a, b = d.jacobi_solve(st, rx, ry), o.jacobi_solve(st, rx, ry)
        assert a["iters"] == b["iters"] == cap
        assert abs(a["error"] - b["error"]) <= SUM_TOL * abs(b["error"])
        np.testing.assert_array_equal(d.get_field("u"), o.get_field("u"))
        np.testing.assert_array_equal(d.get_field("r"), o.get_field("r"))
        d.close(); o.close()
    from tealeaf_jl_b200 import lib
    d, _ = tl.initialiseapp(classic_settings(16, steps=1), backend=_device())
    with pytest.raises(lib.TeaLeafError):
        d.jacobi_init(0, 1.0, 1.0)          # Jacobi.jl:34-36
    d.close()


@pytest.mark.parametrize("solver", ["cg", "cheby", "ppcg", "jacobi"])
@pytest.mark.parametrize("hd,coef", [(3, 1), (4, 2)])
def test_fused_solves_with_other_halo_depths_and_coefficient(solver, hd, coef):
    """halo_depth != 2 changes the padded layout (pitch, row offsets) and the range kx/ky are set on;
    coefficient = 2 is RECIP_CONDUCTIVITY (src/settings.jl:14)."""
    s = lambda: classic_settings(90, ny=70, steps=1, solver=solver, halodepth=hd, coefficient=coef,
                                 maxiters=400 if solver == "jacobi" else 10000)
    dev, ora = run(_device(), s()), run(_oracle(), s())
    assert_parity(dev, ora, iter_slack=1 if solver == "cg" else 0, fields=("u", "energy", "kx", "ky"))


def test_errorswitch_and_maxiters_paths():
    for solver in ("cheby", "ppcg"):
        s = lambda: classic_settings(96, steps=1, solver=solver, errorswitch=True, epslim=1e-3)
        assert_parity(run(_device(), s()), run(_oracle(), s()), iter_slack=0)
    for solver, cap in (("cg", 17), ("cheby", 45), ("ppcg", 33)):
        s = lambda: classic_settings(96, steps=1, solver=solver, maxiters=cap)
        dev, ora = run(_device(), s()), run(_oracle(), s())
        assert dev[1][0]["iters"] == ora[1][0]["iters"] == cap
        assert rel(dev[0].get_field("u"), ora[0].get_field("u")) < 1e-9


def test_fused_and_stepwise_device_paths_agree():
    s = lambda: classic_settings(150, ny=90, steps=1, solver="cg")
    a, b = run(_device(), s()), run(_device(), s(), stepwise=True)
    assert abs(a[1][0]["iters"] - b[1][0]["iters"]) <= 1
    assert rel(a[0].get_field("u"), b[0].get_field("u")) < 1e-11


def test_cg_coefficients_are_returned():
    s = classic_settings(64, steps=1)
    d, _ = tl.initialiseapp(s, backend=_device())
    o, _ = tl.initialiseapp(classic_settings(64, steps=1), backend=_oracle())
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    for c in (d, o):
        tl.haloupdate(c, s, 1, ["energy", "density"])
    info = d.cg_solve(s, rx, ry)
    o.cg_solve(s, rx, ry)
    n = min(info["iters"], 40)
    np.testing.assert_allclose(d.cgalpha[:n], o.cgalpha[:n], rtol=1e-9)
    np.testing.assert_allclose(d.cgbeta[:n], o.cgbeta[:n], rtol=1e-9)


def test_run_to_run_determinism():
    outs = []
    for _ in range(2):
        c, recs, final, _ = run(_device(), classic_settings(200, steps=1, solver="cg"))
        outs.append((c.get_field("u"), recs[0]["iters"], final["temp"]))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])   # fixed-order reductions, no FP64 atomics
    assert outs[0][1:] == outs[1][1:]


@pytest.mark.parametrize("nx,ny,solver,over", [
    (64, 64, "cg", {}), (200, 120, "cg", {}), (257, 19, "cg", {}), (1, 40, "cg", {}), (3, 3, "cg", {}),
    (700, 523, "cg", {"maxiters": 500}), (2048, 1536, "cg", {"maxiters": 150}),     # more work items than resident CTAs
    (128, 128, "cheby", {}), (96, 160, "ppcg", {"ppcginnersteps": 6})])              # CG presteps of the other solvers
def test_persistent_cg_is_bit_identical(nx, ny, solver, over):
    """Option cg_persist: the CG loop as ONE persistent cooperative kernel (tl_kernels_persist.cuh).
    Same work items, same per-item block sums, same fixed summation order: iteration counts,
    alpha/beta histories and every field are bit-identical to the two-kernels-per-iteration path."""
    outs = []
    for persist in (0, 1):
        s = classic_settings(nx, ny=ny, steps=2, solver=solver, **over)
        chunk, geom = tl.initialiseapp(s, backend=_device())
        chunk.set_option("cg_persist", persist)
        recs, final = tl.diffuse(chunk, s, geom)
        outs.append(([r["iters"] for r in recs], [r["error"] for r in recs], final["temp"],
                     {f: chunk.get_field(f) for f in ("u", "energy", "p", "r", "w")},
                     chunk.cgalpha.copy(), chunk.cgbeta.copy(), [r["kernel_launches"] for r in recs]))
        chunk.close()
    a, b = outs
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2], (a[:3], b[:3])
    for f in a[3]:
        np.testing.assert_array_equal(a[3][f], b[3][f], err_msg=f)
    np.testing.assert_array_equal(a[4], b[4])
    np.testing.assert_array_equal(a[5], b[5])
    if solver == "cg":
        assert max(b[6]) < min(a[6])          # one launch for the whole loop instead of two per iteration


@pytest.mark.parametrize("slots", [3, 4])
@pytest.mark.parametrize("nx,ny,solver,over", [(200, 120, "cg", {}), (257, 19, "cg", {}), (1, 40, "cg", {}), (63, 64, "cg", {}), (64, 5, "cg", {}),
                                              (700, 523, "cg", {"maxiters": 400}), (2048, 1536, "cg", {"maxiters": 120}),
                                              (96, 160, "ppcg", {"ppcginnersteps": 6}), (129, 70, "cg", {"halodepth": 3})])
def test_kernel_a_tma_is_bit_identical(nx, ny, solver, over, slots):
    """Option a_tma: kernel A with its row ring filled by TMA (cp.async.bulk.tensor 68 x 1 boxes, mbarrier per ring slot,
    one elected lane issues; tl_kernels_tma.cuh) -- same rows, same arithmetic, same summation order as the cp.async ring."""
    outs = []
    for d in (0, slots):
        s = classic_settings(nx, ny=ny, steps=2, solver=solver, **over)
        chunk, geom = tl.initialiseapp(s, backend=_device())
        chunk.set_option("a_tma", d)
        recs, final = tl.diffuse(chunk, s, geom)
        outs.append(([r["iters"] for r in recs], [r["error"] for r in recs], final["temp"],
                     {f: chunk.get_field(f) for f in ("u", "energy", "p", "r", "w")}, chunk.cgalpha.copy()))
        chunk.close()
    a, b = outs
    assert a[:3] == b[:3], (a[:3], b[:3])
    for f in a[3]:
        np.testing.assert_array_equal(a[3][f], b[3][f], err_msg=f)
    np.testing.assert_array_equal(a[4], b[4])


LAZY_U_CASES = [(200, 120, "cg", {}, {}), (257, 19, "cg", {}, {}), (1, 40, "cg", {}, {}), (64, 5, "cg", {}, {"use_graph": 0}),
                (96, 80, "cg", {"maxiters": 1}, {}), (96, 80, "cg", {"maxiters": 2}, {}), (96, 80, "cg", {"maxiters": 3}, {}),
                (96, 80, "cg", {"maxiters": 4}, {"graph_iters": 3}), (96, 80, "cg", {"maxiters": 9}, {"graph_iters": 1}),
                (700, 523, "cg", {"maxiters": 400}, {}), (700, 523, "cg", {"maxiters": 401}, {"ring_stages": 4}),
                (300, 260, "cg", {"maxiters": 77}, {"ring_stages": 6}), (200, 120, "cg", {"maxiters": 33}, {"graph_iters": 2}),
                (200, 121, "cg", {"maxiters": 50}, {"graph_iters": 6, "use_graph": 0}), (333, 222, "cg", {}, {"cg_lazy_heavy_ctas": 3}), (2048, 1536, "cg", {"maxiters": 121}, {}),
                (129, 70, "cg", {"halodepth": 3}, {}), (128, 96, "cheby", {}, {}), (130, 97, "cheby", {"presteps": 31}, {}),
                (96, 160, "ppcg", {"ppcginnersteps": 6}, {}), (96, 161, "ppcg", {"ppcginnersteps": 5, "presteps": 29}, {})]


@pytest.mark.parametrize("nx,ny,solver,over,opts", LAZY_U_CASES,
                         ids=[f"{c[2]}-{c[0]}x{c[1]}-" + "-".join(f"{k}{v}" for k, v in {**c[3], **c[4]}.items()) for c in LAZY_U_CASES])
def test_lazy_u_update_is_bit_identical(nx, ny, solver, over, opts):
    """Option cg_lazy_u (default on): kernel A of the CG loop advances u every second launch with both pending updates,
    u = (u + alpha(it-2) p(it-2)) + alpha(it-1) p(it-1), and the flush applies the one or two that remain -- the same
    operations in the same order as an update per iteration (CG.jl:95): every field and scalar has the same bits, for odd
    and even iteration counts, odd graph lengths, every ring depth, and across a switch to Chebyshev / PPCG."""
    outs = []
    for lazy in (0, 1):
        s = classic_settings(nx, ny=ny, steps=2, solver=solver, **over)
        chunk, geom = tl.initialiseapp(s, backend=_device())
        chunk.set_option("cg_lazy_u", lazy)
        for k, v in opts.items():
            chunk.set_option(k, v)
        # chunks of an odd number of iterations cannot alternate the two kernels by position: an update per launch
        assert chunk.get_option("cg_u_mode") == (2 if lazy and opts.get("graph_iters", 8) % 2 == 0 else 1)
        recs, final = tl.diffuse(chunk, s, geom)
        outs.append(([(r["iters"], r["cg_iters"], r["error"]) for r in recs], final["temp"],
                     {f: chunk.get_field(f) for f in ("u", "energy", "p", "r", "w", "sd")}, chunk.cgalpha.copy(), chunk.cgbeta.copy()))
        chunk.close()
    a, b = outs
    assert a[:2] == b[:2], (a[:2], b[:2])
    for f in a[2]:
        np.testing.assert_array_equal(a[2][f], b[2][f], err_msg=f)
    np.testing.assert_array_equal(a[3], b[3])
    np.testing.assert_array_equal(a[4], b[4])


@pytest.mark.parametrize("depth", [6, 8])
@pytest.mark.parametrize("nx,ny", [(200, 120), (257, 19), (1, 40), (700, 523)])
def test_kernel_b_ring_is_bit_identical(nx, ny, depth):
    """Option b_ring: kernel B with the cp.async ring instead of register batches -- same rows, same
    arithmetic, same summation order."""
    outs = []
    for d in (0, depth):
        s = classic_settings(nx, ny=ny, steps=1, solver="cg", maxiters=400)
        chunk, geom = tl.initialiseapp(s, backend=_device())
        chunk.set_option("b_ring", d)
        recs, final = tl.diffuse(chunk, s, geom)
        outs.append((recs[0]["iters"], recs[0]["error"], final["temp"], chunk.get_field("u"), chunk.get_field("r"),
                     chunk.cgalpha.copy()))
        chunk.close()
    a, b = outs
    assert a[:3] == b[:3]
    for x, y in zip(a[3:], b[3:]):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("nx,ny,over", [(128, 128, {}), (96, 160, {}), (65, 70, {}), (257, 19, {}), (61, 300, {}),
                                        (700, 523, {"maxiters": 900}), (1500, 1100, {"maxiters": 1100}),
                                        (128, 128, {"maxiters": 77}), (128, 128, {"maxiters": 78})])   # last permitted iteration: first / second half of a pair
def test_cheby_pair_is_bit_identical(nx, ny, over):
    """Option cheby_pair: two Chebyshev iterations per pass (k_cheby_pair_ring, temporal blocking).  The
    intermediate u', p' never reach memory and are recomputed redundantly at warp-task borders with
    the same expressions on the same inputs: every field is bit-identical to one kernel per iteration."""
    outs = []
    for pair in (0, 1):
        s = classic_settings(nx, ny=ny, steps=2, solver="cheby", **over)
        chunk, geom = tl.initialiseapp(s, backend=_device())
        chunk.set_option("cheby_pair", pair)
        recs, final = tl.diffuse(chunk, s, geom)
        outs.append(([(r["iters"], r["cg_iters"], r["cheby_iters"], r["est_iters"]) for r in recs], [r["error"] for r in recs],
                     final["temp"], {f: chunk.get_field(f) for f in ("u", "energy", "p", "w", "r")},
                     [r["kernel_launches"] for r in recs]))
        chunk.close()
    a, b = outs
    assert a[0] == b[0], (a[0], b[0])
    for f in a[3]:
        np.testing.assert_array_equal(a[3][f], b[3][f], err_msg=f)
    assert a[2] == b[2]
    for ea, eb in zip(a[1], b[1]):               # the norm is summed over different warp tasks
        assert abs(ea - eb) <= 1e-10 * abs(ea)
    if max(r[2] for r in a[0]) > 20:
        assert sum(b[4]) < sum(a[4])             # fewer launches: two iterations per kernel


@pytest.mark.parametrize("nx,ny,inner,over", [(128, 128, 10, {}), (96, 160, 4, {}), (65, 70, 6, {}), (257, 19, 2, {}), (61, 300, 10, {}),
                                              (128, 128, 7, {}), (96, 160, 3, {}), (257, 19, 5, {}),      # odd: pairs + one trailing single step
                                              (700, 523, 10, {"maxiters": 700}), (1500, 1100, 8, {"maxiters": 900}),
                                              (1500, 1100, 9, {"maxiters": 900})])
def test_ppcg_pair_is_bit_identical(nx, ny, inner, over):
    """Option ppcg_pair: two PPCG inner steps per pass (k_ppcg_pair_ring).  Per-cell arithmetic is identical
    to one kernel per inner step; the outer iteration's sum(r.r) is added up over different warp tasks and
    feeds beta, so whole solves agree to the rounding of that sum (iteration counts identical, fields
    1e-11 of max|u|), not bit for bit."""
    outs = []
    for pair in (0, 1):
        s = classic_settings(nx, ny=ny, steps=2, solver="ppcg", ppcginnersteps=inner, **over)
        chunk, geom = tl.initialiseapp(s, backend=_device())
        chunk.set_option("ppcg_pair", pair)
        recs, final = tl.diffuse(chunk, s, geom)
        outs.append(([(r["iters"], r["cg_iters"], r["cheby_iters"], r["inner_total"]) for r in recs], [r["error"] for r in recs],
                     final["temp"], {f: chunk.get_field(f) for f in ("u", "energy", "p", "sd", "w")},
                     [r["kernel_launches"] for r in recs]))
        chunk.close()
    a, b = outs
    assert a[0] == b[0], (a[0], b[0])
    for ea, eb in zip(a[1], b[1]):               # a converged residual norm is rounding noise of the sums: observed 2e-8
        assert abs(ea - eb) <= 1e-5 * abs(ea)
    # the outer recurrence consumes the (differently summed) norm, so fields agree to rounding of the sums, not bitwise
    scale = np.abs(a[3]["u"]).max()
    for f in a[3]:
        assert np.abs(a[3][f] - b[3][f]).max() <= 1e-11 * scale, f
    assert abs(a[2] - b[2]) <= 1e-12 * abs(a[2])


def test_persistent_cg_matches_oracle_through_env_default(monkeypatch):
    monkeypatch.setenv("TEALEAF_B200_OPTS", "cg_persist=1")
    s = lambda: classic_settings(150, ny=90, steps=2, solver="cg")
    dev = run(_device(), s())
    ora = run(_oracle(), s())
    assert_parity(dev, ora, fields=("u", "energy", "u0", "kx", "ky"), aux=("p", "w"))
    monkeypatch.setenv("TEALEAF_B200_OPTS", "no_such_option=1")
    with pytest.raises(Exception):
        tl.initialiseapp(s(), backend=_device())


@pytest.mark.parametrize("idx", range(12))
def test_golden(idx):
    with open(GOLDEN) as fh:
        case = json.load(fh)["cases"][idx]
    s = classic_settings(case["nx"], ny=case["ny"], steps=case["steps"], solver=case["solver"])
    chunk, recs, final, summaries = run(_device(), s)
    slack = 1 if case["solver"] == "cg" else 0
    for r, want in zip(recs, case["iters"]):
        assert abs(r["iters"] - want) <= slack, (r, want)
    for got, want in zip(summaries, case["summaries"]):
        for k, v in zip(("vol", "mass", "ie", "temp"), got):
            assert abs(v - want[k]) <= SUMMARY_TOL * abs(want[k]), (k, v, want[k])
    u = chunk.get_field("u")[2:-2, 2:-2]
    probe = case["u_probe"]
    assert abs(u.sum() - probe["sum"]) <= 1e-10 * abs(probe["sum"])
    assert abs(u.max() - probe["max"]) <= 1e-9 * abs(probe["max"])
    assert abs(u[case["nx"] // 2, case["ny"] // 2] - probe["center"]) <= 1e-9 * abs(probe["max"])


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json sizes; the serial oracle would take minutes here)
# ---------------------------------------------------------------------------------------------
def test_full_size_4096_properties():
    s = classic_settings(4096, steps=1, solver="cg")
    chunk, geom = tl.initialiseapp(s, backend=_device())
    recs, final = tl.diffuse(chunk, s, geom)
    it = recs[0]["iters"]
    assert 4000 < it < 5200            # ~1.1 N iterations (SURVEY Appendix C extrapolation)
    u = chunk.get_field("u")[2:-2, 2:-2]
    u0 = chunk.get_field("u0")[2:-2, 2:-2]
    assert abs(u.sum() - u0.sum()) <= 1e-11 * abs(u0.sum())      # energy conservation
    assert np.abs(chunk.get_field("r")).max() < 1e-9             # true residual after residual!
    assert u.min() > 0
    # halo of u reflects its interior (haloupdate! {u,p} at CG.jl:22)
    full = chunk.get_field("u")
    np.testing.assert_array_equal(full[1, 2:-2], full[2, 2:-2])
    np.testing.assert_array_equal(full[2:-2, -2], full[2:-2, -3])
