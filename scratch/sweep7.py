# PDL on/off: us per sweep for CG / Chebyshev / PPCG at several sizes, single GPU
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
from tealeaf_jl_b200.solvers import get_solver
for N, mi in ((256, 600), (1024, 1500), (4096, 1400)):
    for solver in ('cg', 'cheby', 'ppcg'):
        s = classic_settings(N, steps=1, solver=solver, maxiters=mi)
        chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
        rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
        res = []
        for pdl in (0, 1, 0, 1):
            chunk.set_option('use_pdl', pdl)
            best = None
            for _ in range(3):
                chunk.copy_field("energy", "energy0")
                tl.haloupdate(chunk, s, 1, ["energy", "density"])
                info = get_solver(solver).solve(chunk, s, rx, ry)
                sweeps = info['cg_iters'] + info['cheby_iters'] + info['inner_total']
                t = 1e3 * info['solve_ms'] / sweeps
                best = t if best is None else min(best, t)
            res.append(best)
        print(f"N={N} {solver:5s} iters={info['iters']} sweeps={sweeps}: us/sweep pdl0={res[0]:.2f} pdl1={res[1]:.2f} pdl0={res[2]:.2f} pdl1={res[3]:.2f}", flush=True)
        chunk.close()
