#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
out=gpurun_out/s6_config_bench.jsonl
: > $out
python tests/mgpu_check.py > /dev/null 2>&1
$TR tests/mgpu_check.py 2>&1 | grep mgpu | tee gpurun_out/s6_mgpu.log
TEALEAF_GRID=2x1 $TR tests/mgpu_check.py cg cheby 2>&1 | grep mgpu | tee -a gpurun_out/s6_mgpu.log
$TR tools/config_bench.py --solver cg --global 4096 --max-iters 1000 >> $out 2>gpurun_out/s6.err
$TR tools/config_bench.py --solver cheby --global 4096 --max-iters 2000 >> $out 2>>gpurun_out/s6.err
$TR tools/config_bench.py --solver cg --global 1024 --max-iters 2000 >> $out 2>>gpurun_out/s6.err
TEALEAF_GRID=2x1 $TR tools/config_bench.py --solver cg --global 4096 --max-iters 1000 --comm fused >> $out 2>>gpurun_out/s6.err
python - >> $out 2>>gpurun_out/s6.err <<'PY'
# the half tiles alone on one GPU: the per-tile floor of the 2-GPU runs above
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
for nx, ny, mi in ((4096, 2048, 1000), (2048, 4096, 1000), (1024, 512, 1000)):
    s = classic_settings(nx, ny=ny, steps=1, solver='cg', maxiters=mi)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    best = None
    for _ in range(3):
        chunk.copy_field("energy", "energy0")
        info = chunk.cg_solve(s, rx, ry)
        best = info if best is None or info['solve_ms'] < best['solve_ms'] else best
    ka = chunk.time_kernel('cg_fused_w', 30); kb = chunk.time_kernel('cg_fused_r', 30)
    print(json.dumps({"single_tile": [nx, ny], "iters": best['iters'], "us_per_iter": 1e3 * best['solve_ms'] / best['iters'], "A_us": ka * 1e3, "B_us": kb * 1e3}))
    chunk.close()
PY
cat $out; tail -3 gpurun_out/s6.err
