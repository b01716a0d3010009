"""tl_create_multi: ONE context (one `Chunk`, run.jl:45-47) spread over several GPUs of one process.  Here the tiles
share cuda:0 (`devices` repeats index 0 -- the test mode for single-GPU boxes; tests/test_multi_gpu.py repeats the
solves across real GPUs).  The host code is exactly the single-chunk code: `initialiseapp`, `diffuse`, `get_field`
know nothing about tiles.  Bars as everywhere: iteration counts exact (CG +-1), summaries 1e-10, u / energy 1e-9."""
import re
import subprocess

import numpy as np
import pytest

import tealeaf_jl_b200 as tl
from conftest import classic_settings
from test_gpu_parity import _oracle, assert_parity, run

pytestmark = pytest.mark.gpu


def multi_backend(ngpus, px=0, py=0, devices=None):
    from tealeaf_jl_b200.device import DeviceChunk
    devices = [0] * ngpus if devices is None else devices
    return lambda *a, **k: DeviceChunk.multi(*a, ngpus=ngpus, devices=devices, px=px, py=py, **k)


CASES = [("cg", 4, 0, 0, 200, 150, {}), ("cg", 2, 0, 0, 130, 77, {}), ("cheby", 4, 0, 0, 129, 140, {}), ("cheby", 3, 3, 1, 200, 64, {}),
         ("ppcg", 4, 0, 0, 131, 150, {"ppcginnersteps": 5}), ("ppcg", 6, 2, 3, 160, 190, {"ppcginnersteps": 10}),
         ("jacobi", 4, 0, 0, 160, 130, {"maxiters": 120}), ("cg", 1, 0, 0, 64, 48, {})]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("solver,n,px,py,nx,ny,over", CASES, ids=[f"{c[0]}-{c[1]}tiles-{c[4]}x{c[5]}" for c in CASES])
def test_multi_context_solves_match_oracle(solver, n, px, py, nx, ny, over):
    s = lambda: classic_settings(nx, ny=ny, steps=2, solver=solver, **over)
    dev = run(multi_backend(n, px, py), s())
    ora = run(_oracle(), s())
    assert_parity(dev, ora, iter_slack=1 if solver == "cg" else 0)
    # the gathered global arrays carry the physical halos too (reflected u)
    if solver == "cg":       # CG keeps the halo of u reflected (haloupdate!(u, p), CG.jl:22)
        u = dev[0].get_field("u")
        np.testing.assert_array_equal(u[1, 2:-2], u[2, 2:-2])
        np.testing.assert_array_equal(u[2:-2, -2], u[2:-2, -3])
    dev[0].close()


def test_multi_context_scatter_gather_roundtrip_and_painter():
    from tealeaf_jl_b200.chunk import HostGeometry, paint_states
    rng = np.random.default_rng(5)
    for n, px, py, nx, ny, hd in ((4, 2, 2, 37, 21, 2), (6, 3, 2, 50, 33, 3), (3, 1, 3, 16, 40, 2)):
        s = classic_settings(nx, ny=ny, steps=1, halodepth=hd)
        c = multi_backend(n, px, py)(nx, ny, hd, 100)
        a = np.asfortranarray(rng.standard_normal((nx + 2 * hd, ny + 2 * hd)))
        c.set_field("sd", a)
        np.testing.assert_array_equal(c.get_field("sd"), a)              # every element comes back from exactly one tile
        geom = HostGeometry(s)
        c.paint_states(s, geom)                                           # every tile paints at its own offset
        density, energy0, u = paint_states(s, geom)
        np.testing.assert_array_equal(c.get_field("density"), density)
        np.testing.assert_array_equal(c.get_field("energy0"), energy0)
        hdx = slice(hd, -hd)
        np.testing.assert_array_equal(c.get_field("u")[hdx, hdx], u[hdx, hdx])
        # an all-tiles scalar through the per-function ABI
        assert abs(c.norm2("sd") - (a[hdx, hdx] ** 2).sum()) <= 1e-12 * (a[hdx, hdx] ** 2).sum()
        vol, mass, ie, temp = c.fieldsummary(0.5)
        assert abs(vol - 0.5 * nx * ny) < 1e-9 and abs(mass - 0.5 * density[hdx, hdx].sum()) <= 1e-12 * mass
        c.close()


def test_multi_context_stepwise_kernels_and_options():
    """The reference's own control flow, kernel by kernel (CG.jl:7-44), on a multi context."""
    s = lambda: classic_settings(96, ny=80, steps=1, solver="cg")
    dev = run(multi_backend(4), s(), stepwise=True)
    ora = run(_oracle(), s(), stepwise=True)
    assert_parity(dev, ora, iter_slack=1)
    c = dev[0]
    c.set_option("graph_iters", 3)
    assert c.get_option("graph_iters") == 3 and c.launch_count() > 0
    from tealeaf_jl_b200.lib import TeaLeafError
    with pytest.raises(TeaLeafError):
        c.comm_export()
    c.close()


def test_multi_context_argument_errors():
    from tealeaf_jl_b200.device import DeviceChunk
    from tealeaf_jl_b200.lib import TeaLeafError
    with pytest.raises(TeaLeafError):
        DeviceChunk.multi(16, 16, 2, 10, ngpus=4, devices=[0] * 4, px=3, py=2)      # px*py != ngpus
    with pytest.raises(TeaLeafError):
        DeviceChunk.multi(16, 16, 2, 10, ngpus=2, devices=[0, 99])                  # no such device


@pytest.mark.timeout(600)
@pytest.mark.parametrize("solver,extra", [("cg", []), ("ppcg", ["10000", "6"])])
def test_c_host_with_one_context_over_four_tiles(tmp_path, solver, extra):
    """examples/tea_c_host.c --gpus 4: the same C loop, one tl_ctx, four tiles."""
    from test_c_host import build_c_host, parse
    exe = build_c_host(tmp_path)
    nx, ny = 150, 131
    outs = {}
    for label, flags in (("single", []), ("multi", ["--gpus", "4", "--share-gpu"])):
        r = subprocess.run([exe, *flags, str(nx), str(ny), "2", solver, *extra], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[label] = (parse(r.stdout), r.stdout)
    (s1, sum1), (s4, sum4) = outs["single"][0], outs["multi"][0]
    slack = 1 if solver == "cg" else 0
    assert all(abs(a[0] - b[0]) <= slack for a, b in zip(s1, s4)), (s1, s4)
    assert all(abs(a / b - 1) < 1e-10 for a, b in zip(sum1, sum4)), (sum1, sum4)
    usum = float(re.search(r"usum (\S+)", outs["multi"][1]).group(1))
    from tealeaf_jl_b200.device import DeviceChunk
    over = {"maxiters": int(extra[0]), "ppcginnersteps": int(extra[1])} if extra else {}
    st = classic_settings(nx, ny=ny, steps=2, solver=solver, **over)
    chunk, geom = tl.initialiseapp(st, backend=DeviceChunk)
    tl.diffuse(chunk, st, geom)
    want = chunk.get_field("u")[2:-2, 2:-2].sum()
    assert abs(usum / want - 1) < 1e-10
    chunk.close()
