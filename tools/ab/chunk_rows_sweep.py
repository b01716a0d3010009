# Sweep: rows per warp task of the stencil kernels (chunk_rows) with the lazy-u CG loop, whole solves, several tile sizes
import sys, os
sys.path.insert(0, os.getcwd())
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
from tealeaf_jl_b200.solvers import get_solver
sizes = [(2048, 2048, 1500), (3072, 3072, 1000), (4096, 4096, 1000), (4096, 8192, 500), (6144, 6144, 400), (2048, 1024, 1500), (1024, 1024, 1500)]
rows = [-1, 8, 10, 11, 12, 13, 14, 16, 20, -1]
for nx, ny, iters in sizes:
    s = classic_settings(nx, ny=ny, steps=1, solver='cg', maxiters=iters)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    out = []
    for cr in rows:
        chunk.set_option("chunk_rows", cr)
        best = None
        for _ in range(3):
            chunk.copy_field("energy", "energy0")
            tl.haloupdate(chunk, s, 1, ["energy", "density"])
            info = get_solver('cg').solve(chunk, s, rx, ry)
            if best is None or info["solve_ms"] < best["solve_ms"]:
                best = info
        out.append(f"{cr}->{int(chunk.get_option('rows_per_chunk'))} rows (grid {int(chunk.get_option('fused_grid'))}): {1e3 * best['solve_ms'] / best['iters']:.2f}")
    print(f"[chunk_rows] {nx}x{ny} us/iteration: " + "; ".join(out), flush=True)
    chunk.close()
