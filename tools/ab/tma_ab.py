# A/B: CG kernel A with the cp.async ring (default) vs the TMA ring (option a_tma = 3 / 4 row slots), kernels alone and whole CG solves
import sys, os
sys.path.insert(0, os.getcwd())
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
from tealeaf_jl_b200.solvers import get_solver
for N, iters in ((4096, 1500), (8192, 600), (2048, 1500)):
    s = classic_settings(N, steps=1, solver='cg', maxiters=iters)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    for opts in ({"a_tma": 0}, {"a_tma": 0, "ring_stages": 3}, {"a_tma": 0, "ring_stages": 4}, {"a_tma": 4}, {"a_tma": 3}):
        chunk.set_option("ring_stages", -1)
        for k, v in opts.items():
            chunk.set_option(k, v)
        chunk.cg_init(s.coefficient, rx, ry)
        ka = min(chunk.time_kernel('cg_fused_w', 30) for _ in range(3))
        best = None
        for _ in range(2):
            chunk.copy_field("energy", "energy0")
            tl.haloupdate(chunk, s, 1, ["energy", "density"])
            info = get_solver('cg').solve(chunk, s, rx, ry)
            if best is None or info["solve_ms"] < best["solve_ms"]:
                best = info
        print(f"[tma_ab] N={N} {opts}: kernel A {ka * 1e3:.1f} us ({64 * N * N / ka / 1e9:.2f} TB/s physical); "
              f"solve {best['iters']} iterations {best['solve_ms']:.1f} ms = {1e3 * best['solve_ms'] / best['iters']:.1f} us/iteration, error {best['error']!r}", flush=True)
    chunk.close()
