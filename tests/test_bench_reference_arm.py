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


def test_recorded_bench_lines_carry_the_contract_keys():
    """The B200 arm's lines recorded at HEAD (profiles/r02q_bench_n{1,2,4}_*.json, written by bench.py on the GPU boxes):
    every key the measurement contract names is there and self-consistent."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02q_bench_n[124]_*.json")))
    assert len(files) >= 3, files
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
            assert k in d, (f, k)
        assert d["metric"] == "solver cell-iterations/sec" and d["dtype"] == "f64" and d["scaling"] == "weak" and d["warmup"] >= 3
        assert "workload" in d["config"] and "model" not in d["config"]
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert r["traffic"] is None or r["traffic"] > 0
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] <= 1.02 * d["value"]
        assert d["gpu_launches"] > 0
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
        if d["n_gpus"] == 1:
            assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
            assert "weak_tile_16384" in d["other_configs"] and "chebyshev_4096" in d["other_configs"] and "ppcg_8192" in d["other_configs"]
        else:
            assert "same_tile_single_gpu" in d and "chebyshev_4096" in d["other_configs"] and "ppcg_8192" in d["other_configs"]
