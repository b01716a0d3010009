# A/B of the balanced (two CTAs per SM) chunking on mid-size tiles: whole solves, graph-launched
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
from tealeaf_jl_b200.solvers import get_solver
shapes = [(2048, 1024), (1024, 1024), (2048, 2048), (1024, 512), (512, 512), (1536, 1024), (4096, 1024), (4096, 4096)]
for nx, ny in shapes:
    for solver, iters in (("cg", 600), ("cheby", 1500), ("ppcg", 1200)):
        if nx * ny >= 4096 * 4096 and solver != "cg":
            continue
        out = {}
        for bal in (0, 1, 0, 1):
            s = classic_settings(nx, ny=ny, steps=1, solver=solver, maxiters=iters)
            chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
            chunk.set_option("balanced_tiling", bal)
            rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
            best = None
            for _ in range(3):
                chunk.copy_field("energy", "energy0")
                tl.haloupdate(chunk, s, 1, ["energy", "density"])
                info = get_solver(solver).solve(chunk, s, rx, ry)
                if best is None or info["solve_ms"] < best["solve_ms"]:
                    best = info
            sweeps = best["cg_iters"] + best["cheby_iters"] + best["inner_total"]
            out.setdefault(bal, []).append((best["solve_ms"], sweeps, best["iters"], best["error"]))
            chunk.close()
        a, b = min(out[0]), min(out[1])
        print(f"[bal] {nx}x{ny} {solver:5s} sweeps={a[1]} (cg/outer iters {a[2]})  old {1e3 * a[0] / a[1]:.2f} us/sweep  balanced {1e3 * b[0] / b[1]:.2f} us/sweep  "
              f"({100 * (a[0] / b[0] - 1):+.1f} %)  err rel diff {abs(a[3] - b[3]) / max(abs(a[3]), 1e-300):.1e}", flush=True)
