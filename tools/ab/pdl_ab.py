# A/B: programmatic dependent launch released before the kernel tails (option use_pdl) + the kernel-boundary micro-profile
import sys, os
sys.path.insert(0, os.getcwd())
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
from tealeaf_jl_b200.solvers import get_solver
for N, cap in ((1024, 1200), (2048, 1500), (4096, 1500)):
    for solver in ("cg", "cheby", "ppcg"):
        s = classic_settings(N, steps=1, solver=solver, maxiters=cap if solver == "cg" else cap + 600)
        chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
        rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
        for pdl in (0, 1, 0, 1):
            chunk.set_option("use_pdl", pdl)
            best = None
            for _ in range(2):
                chunk.copy_field("energy", "energy0")
                tl.haloupdate(chunk, s, 1, ["energy", "density"])
                info = get_solver(solver).solve(chunk, s, rx, ry)
                if best is None or info["solve_ms"] < best["solve_ms"]:
                    best = info
            chunk.set_option("prof", 1)
            chunk.copy_field("energy", "energy0")
            tl.haloupdate(chunk, s, 1, ["energy", "density"])
            get_solver(solver).solve(chunk, s, rx, ry)
            prof = {k: round(chunk.get_option("prof_" + k), 2) for k in ("body_us", "sum_us", "xchg_us", "gap_us", "kernels")}
            chunk.set_option("prof", 0)
            sweeps = best["iters"] + best["inner_total"]
            print(f"[pdl_ab] N={N} {solver} use_pdl={pdl}: {best['solve_ms']:.1f} ms, {1e3 * best['solve_ms'] / sweeps:.2f} us/sweep "
                  f"(iters {best['iters']}, inner {best['inner_total']}, error {best['error']!r}); boundary profile per kernel {prof}", flush=True)
        chunk.close()
