"""bench.py --impl reference on the CPU (no GPU needed): the line's schema, that the arm never loads the CUDA library
(TEALEAF_B200_LIB points at a file that does not exist: loading it would raise), that nothing else reaches stdout, and
that under torchrun the ranks other than 0 exit 0 without work."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_arm(extra_env, *flags):
    env = dict(os.environ, TEALEAF_B200_LIB="/nonexistent/libtealeaf_b200.so", **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *flags],
                          capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_line():
    r = run_arm({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout              # ONE JSON line and nothing else on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "solver cell-iterations/sec" and d["unit"] == "cell-iterations/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 1e7 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline_serial"]["cores"] == 1 and 0 < d["cpu_baseline_serial"]["value"] <= 1.5 * d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "4096x4096" in d["config"]["sample"] and d["config"]["sample"] == d["cpu_baseline"]["sample"]
    assert d["vs_baseline"] is None


def test_reference_arm_other_ranks_do_nothing():
    r = run_arm({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"}, "--gpus", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""
