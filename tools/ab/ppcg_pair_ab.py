# A/B: PPCG inner steps one kernel per step vs two steps per pass (option ppcg_pair)
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
from tealeaf_jl_b200.solvers import get_solver
for N, iters in ((8192, 0), (2048, 0)):
    s = classic_settings(N, steps=1, solver='ppcg', maxiters=max(iters, 400))
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    k1 = min(chunk.time_kernel('ppcg_inner', 30) for _ in range(3))
    k2 = min(chunk.time_kernel('ppcg_pair', 30) for _ in range(3))
    print(f"[ppair] N={N} inner step kernel {k1 * 1e3:.1f} us/step ({64 * N * N / k1 / 1e9:.2f} TB/s physical); pair kernel {k2 * 1e3:.1f} us per launch = "
          f"{k2 * 5e2:.1f} us/step ({100 * (2 * k1 / k2 - 1):+.1f} %)", flush=True)
    if iters:
        res = {}
        for pair in (0, 1):
            chunk.set_option("ppcg_pair", pair)
            best = None
            for _ in range(2):
                chunk.copy_field("energy", "energy0")
                tl.haloupdate(chunk, s, 1, ["energy", "density"])
                info = get_solver('ppcg').solve(chunk, s, rx, ry)
                if best is None or info["solve_ms"] < best["solve_ms"]:
                    best = info
            res[pair] = best
            print(f"[ppair] N={N} solve ppcg_pair={pair}: {best['solve_ms']:.1f} ms  cg {best['cg_iters']} + outer {best['cheby_iters']} x inner {best['inner_total']}  error {best['error']!r}  launches {best['kernel_launches']}", flush=True)
        a, b = res[0], res[1]
        if a["inner_total"]:
            print(f"[ppair] N={N} PPCG phase: {1e3 * (a['solve_ms'] - b['solve_ms']) / a['inner_total']:.1f} us saved per inner step of ~{k1 * 1e3:.1f}", flush=True)
    chunk.close()
