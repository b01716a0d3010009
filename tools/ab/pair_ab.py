# A/B: Chebyshev one kernel per iteration vs two iterations per pass (option cheby_pair)
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
from tealeaf_jl_b200.solvers import get_solver
for N, iters in ((4096, 2600), (8192, 3300), (2048, 1800)):
    s = classic_settings(N, steps=1, solver='cheby', maxiters=iters)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    k1 = min(chunk.time_kernel('cheby_fused', 30) for _ in range(3))
    print(f"[pair] N={N} single-step kernel {k1 * 1e3:.1f} us/iteration ({56 * N * N / k1 / 1e9:.2f} TB/s physical)", flush=True)
    for rows in (16, 32, 64):
        chunk.set_option("pair_rows", rows)
        k2 = min(chunk.time_kernel('cheby_pair', 30) for _ in range(3))
        print(f"[pair] N={N} pair kernel rows={rows}: {k2 * 1e3:.1f} us per launch = {k2 * 5e2:.1f} us/iteration  ({100 * (2 * k1 / k2 - 1):+.1f} %)", flush=True)
    chunk.set_option("pair_rows", 32)
    res = {}
    for pair in (0, 1):
        chunk.set_option("cheby_pair", pair)
        best = None
        for _ in range(2):
            chunk.copy_field("energy", "energy0")
            tl.haloupdate(chunk, s, 1, ["energy", "density"])
            info = get_solver('cheby').solve(chunk, s, rx, ry)
            if best is None or info["solve_ms"] < best["solve_ms"]:
                best = info
        res[pair] = best
        print(f"[pair] N={N} solve cheby_pair={pair}: {best['solve_ms']:.1f} ms  cg {best['cg_iters']} + cheby {best['cheby_iters']} iterations  error {best['error']!r}  launches {best['kernel_launches']}", flush=True)
    a, b = res[0], res[1]
    if a["cheby_iters"]:
        # CG part is identical in both: attribute the difference to the Chebyshev iterations
        print(f"[pair] N={N} Chebyshev phase: {1e3 * (a['solve_ms'] - b['solve_ms']) / a['cheby_iters']:.1f} us saved per iteration of ~{k1 * 1e3:.1f}", flush=True)
    chunk.close()
