mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_c_host.py tests/test_abi.py -x -q ) > gpurun_out/s23_pytest.log 2>&1
tail -5 gpurun_out/s23_pytest.log
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/s23_tea_bm_5.log
import sys, os, json, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
want = {c["x_cells"]: c["temp"] for c in json.load(open("tests/golden/upstream_qa.json"))["cases"]}
for n in (1000, 4000):
    s = classic_settings(n, steps=10, solver="cg")
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    t0 = time.time()
    recs, final = tl.diffuse(chunk, s, geom)
    print(f"tea_bm {n}^2 x 10 steps on B200: temp {final['temp']!r} upstream {want[n]!r} rel {final['temp'] / want[n] - 1:+.3e} "
          f"iterations {sum(r['iters'] for r in recs)} ({[r['iters'] for r in recs]}) solve {sum(r['solve_ms'] for r in recs):.0f} ms wall {time.time() - t0:.1f} s", flush=True)
    chunk.close()
PY
