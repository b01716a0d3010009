"""examples/tea_c_host.c: the reference's timestep loop written in plain C99 against
include/tealeaf_b200.h.  CPU: the header is valid strict C and the library links from C (and fails
loudly without a GPU).  GPU: the C host reproduces the Python mirror's run and the oracle's."""
import os
import re
import subprocess

import pytest

import tealeaf_jl_b200 as tl
from conftest import ROOT, classic_settings

CSRC = os.path.join(ROOT, "tealeaf.jl_b200", "csrc")


def build_c_host(tmp_path):
    import __graft_entry__
    __graft_entry__.build()
    exe = os.path.join(str(tmp_path), "tea_c_host")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-o", exe, os.path.join(ROOT, "examples", "tea_c_host.c"), "-L", CSRC, "-ltealeaf_b200", f"-Wl,-rpath,{CSRC}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_host_compiles_as_strict_c99_and_refuses_to_run_without_a_gpu(tmp_path):
    exe = build_c_host(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, "16", "16", "1", "cg"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


def parse(out):
    steps = [(int(m.group(1)), float(m.group(2))) for m in re.finditer(r"step \d+ iters (\d+) error (\S+)", out)]
    m = re.search(r"summary vol (\S+) mass (\S+) ie (\S+) temp (\S+)", out)
    return steps, tuple(float(g) for g in m.groups())


@pytest.mark.gpu
@pytest.mark.parametrize("solver,nx,ny,extra", [("cg", 96, 80, []), ("cheby", 128, 96, []), ("ppcg", 96, 80, ["10000", "6"]),
                                                ("jacobi", 48, 40, ["400"])])
def test_c_host_matches_the_python_mirror_and_the_oracle(tmp_path, solver, nx, ny, extra):
    exe = build_c_host(tmp_path)
    r = subprocess.run([exe, str(nx), str(ny), "2", solver, *extra], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    steps, summary = parse(r.stdout)
    over = {}
    if len(extra) > 0:
        over["maxiters"] = int(extra[0])
    if len(extra) > 1:
        over["ppcginnersteps"] = int(extra[1])
    from tealeaf_jl_b200.device import DeviceChunk
    from oracle.oracle import OracleChunk
    for backend, tol, slack in ((DeviceChunk, 1e-13, 0), (OracleChunk, 1e-10, 1 if solver == "cg" else 0)):
        s = classic_settings(nx, ny=ny, steps=2, solver=solver, **over)
        chunk, geom = tl.initialiseapp(s, backend=backend)
        recs, final = tl.diffuse(chunk, s, geom)
        assert len(recs) == len(steps) == 2
        for rec, (iters, _err) in zip(recs, steps):
            assert abs(rec["iters"] - iters) <= slack, (backend.__name__, rec["iters"], iters)
        for name, got in zip(("vol", "mass", "ie", "temp"), summary):
            assert abs(got - final[name]) <= tol * abs(final[name]), (backend.__name__, name, got, final[name])


@pytest.mark.gpu
def test_python_cli_mirror_of_run_jl(tmp_path, capsys):
    """`run.py` = run.jl (same flags, run.jl:5-23): deck from -i, -s / -x / -y overrides, -O dump, plus the tea.out
    report; the final summary equals the oracle's."""
    import json
    from tealeaf_jl_b200 import run as cli
    from oracle.oracle import OracleChunk
    deck = os.path.join(ROOT, "decks", "tea_bm_small.in")
    out, dump = tmp_path / "tea.out", tmp_path / "dump.txt"
    cli.main(["-i", deck, "-s", "cheby", "-x", "128", "-y", "96", "-O", str(dump), "--tea-out", str(out)])
    lines = [json.loads(ln) for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    final = lines[-1]["final_summary"]
    s = tl.parse_settings(deck)
    s.solver, s.xcells, s.ycells = "cheby", 128, 96
    s.recompute_spacing()
    chunk, geom = tl.initialiseapp(s, backend=OracleChunk)
    recs, ofinal = tl.diffuse(chunk, s, geom)
    assert [ln["iters"] for ln in lines[:-1]] == [r["iters"] for r in recs] and lines[0]["cheby_iters"] > 0
    for k in ("vol", "mass", "ie", "temp"):
        assert abs(final[k] - ofinal[k]) <= 1e-10 * abs(ofinal[k]), k
    text = out.read_text()
    assert text.count(" Step ") == 2 and "final:" in text and "solver cheby" in text
    assert dump.read_text().count("density0\n") == 2          # one record per timestep (TeaLeaf.jl:68)
