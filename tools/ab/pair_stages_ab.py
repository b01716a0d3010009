# A/B: cp.async ring depth of the pair kernels (option pair_stages: 4 = default, 5), kernels alone
import sys, os
sys.path.insert(0, os.getcwd())
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
for nx, ny in ((4096, 4096), (8192, 8192), (2048, 1024)):
    s = classic_settings(nx, ny=ny, steps=1, solver='cheby', maxiters=400)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    chunk.cg_init(s.coefficient, rx, ry)
    for stages in (4, 5, 4, 5):
        chunk.set_option("pair_stages", stages)
        kc = min(chunk.time_kernel('cheby_pair', 30) for _ in range(3))
        kp = min(chunk.time_kernel('ppcg_pair', 30) for _ in range(3))
        print(f"[pair_stages] {nx}x{ny} stages={stages}: cheby pair {kc * 5e2:.1f} us/iteration, ppcg pair {kp * 5e2:.1f} us/step", flush=True)
    chunk.close()
