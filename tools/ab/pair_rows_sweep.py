# Sweep of the pair kernels' rows per warp task (option pair_rows): wave quantisation of the pair grid
# (4096^2: 69 strips x 128 chunks = 1104 CTAs on 296 resident slots = 3.73 waves at 32 rows).
import sys, os
sys.path.insert(0, os.getcwd())
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
for nx, ny in ((4096, 4096), (2048, 1024), (4096, 2048), (8192, 8192)):
    s = classic_settings(nx, ny=ny, steps=1, solver='cheby', maxiters=400)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    chunk.cg_init(s.coefficient, rx, ry)
    for rows in (16, 20, 24, 28, 30, 32, 36, 40, 41, 44, 48, 56, 64):
        chunk.set_option("pair_rows", rows)
        grid = chunk.get_option("pair_grid"); rpc = chunk.get_option("pair_rows_per_chunk")
        kc = min(chunk.time_kernel('cheby_pair', 30) for _ in range(3))
        kp = min(chunk.time_kernel('ppcg_pair', 30) for _ in range(3))
        print(f"[pair_rows] {nx}x{ny} rows={rows} (effective {rpc:.0f}) grid={grid:.0f} waves={grid / 296:.2f}: cheby pair {kc * 5e2:.1f} us/iteration, ppcg pair {kp * 5e2:.1f} us/step", flush=True)
    chunk.close()
