"""Tiled (2-D domain decomposition) runs.  The GPU test launches tests/mgpu_check.py under
torchrun when the box has >= 2 GPUs; the CPU test covers the host-side plumbing with gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest

import tealeaf_jl_b200 as tl
from tealeaf_jl_b200 import dist as tld
from conftest import classic_settings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, nproc, *args, timeout=900, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT,
                          env={**os.environ, **(env or {})})


def test_decomposition_helpers():
    assert tld.grid_for(1) == (1, 1) and tld.grid_for(2) == (1, 2) and tld.grid_for(4) == (2, 2) and tld.grid_for(8) == (2, 4)
    assert tld.grid_for(6) == (2, 3)
    for n, parts in ((10, 3), (4096, 4), (7, 7), (130, 4)):
        pieces = [tld.split(n, parts, i) for i in range(parts)]
        assert pieces[0][0] == 0 and sum(p[1] for p in pieces) == n
        for a, b in zip(pieces, pieces[1:]):
            assert a[0] + a[1] == b[0]
    # tiles painted independently assemble to the single-chunk painting, halos included
    s = classic_settings(50, ny=37)
    from tealeaf_jl_b200.chunk import HostGeometry, paint_states
    d, e, _ = paint_states(s, HostGeometry(s))
    for px, py in ((1, 2), (2, 2), (2, 4), (3, 1)):
        dt, et = tld.paint_global_from_tiles(s, px, py)
        np.testing.assert_array_equal(dt, d)
        np.testing.assert_array_equal(et, e)


def test_gloo_world2_host_plumbing():
    r = _torchrun("dist_cpu_worker.py", 2, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "dist_cpu_worker OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("grid", ["default", "x-split", "y-split"])
def test_tiled_solvers_match_oracle(grid):
    """Every solver tiled over all GPUs of the box vs the single-chunk oracle: the default
    decomposition (1x2, 2x2, 2x4) and the two 1-D splits (tiles with neighbours on both sides)."""
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    env = {}
    if grid == "x-split":
        env["TEALEAF_GRID"] = f"{n}x1"
    elif grid == "y-split":
        if n == 2:
            pytest.skip("1x2 is the default decomposition on 2 GPUs")
        env["TEALEAF_GRID"] = f"1x{n}"
    r = _torchrun("mgpu_check.py", n, env=env)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_two_devices_in_one_process():
    """Two independent single-tile contexts on two GPUs of the same process (per-device kernel
    attributes, per-context streams): both give the single-GPU result."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from tealeaf_jl_b200.device import DeviceChunk
    outs = []
    for dev in (1, 0, 1):
        s = classic_settings(200, ny=130, steps=1, solver="ppcg")
        chunk, geom = tl.initialiseapp(s, backend=DeviceChunk, device=dev)
        recs, final = tl.diffuse(chunk, s, geom)
        outs.append((recs[0]["iters"], final["temp"], chunk.get_field("u")))
        chunk.close()
    for o in outs[1:]:
        assert o[0] == outs[0][0] and o[1] == outs[0][1]
        np.testing.assert_array_equal(o[2], outs[0][2])



@pytest.mark.gpu
@pytest.mark.parametrize("solver,over", [("cg", {}), ("cheby", {}), ("ppcg", {"ppcginnersteps": 6})])
def test_one_context_over_all_gpus_of_the_box(solver, over):
    """tl_create_multi on real GPUs: ONE process, ONE context, one tile per GPU (peer access between the devices of the
    process); host code identical to the single-GPU code."""
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    from tealeaf_jl_b200.device import DeviceChunk
    from oracle.oracle import OracleChunk
    from test_gpu_parity import assert_parity, run
    s = lambda: classic_settings(384, ny=320, steps=2, solver=solver, **over)
    dev = run(lambda *a, **k: DeviceChunk.multi(*a, ngpus=n, **k), s())
    ora = run(OracleChunk, s())
    assert_parity(dev, ora, iter_slack=1 if solver == "cg" else 0)
    dev[0].close()
