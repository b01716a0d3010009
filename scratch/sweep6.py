# small-mesh tiling sweep: CG us/iter vs chunk_rows / pw_chunk_rows
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
for N in (256, 512, 1024, 2048):
    s = classic_settings(N, steps=1, solver='cg', maxiters=600)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    def solve():
        chunk.copy_field("energy", "energy0")
        best = None
        for _ in range(3):
            info = chunk.cg_solve(s, rx, ry)
            m_ = info['solve_ms'] / info['iters']
            best = m_ if best is None else min(best, m_)
        return best * 1e3
    for cr in (1, 2, 4, 8):
        row = []
        for pcr in (1, 2, 4, 8, 16):
            chunk.set_option('chunk_rows', cr); chunk.set_option('pw_chunk_rows', pcr)
            row.append(solve())
        print(f"N={N} chunk_rows={cr}: us/iter for pw_chunk_rows 1,2,4,8,16 = " + " ".join(f"{v:6.1f}" for v in row), flush=True)
    for gi in (8, 32):
        chunk.set_option('chunk_rows', -1); chunk.set_option('pw_chunk_rows', -1); chunk.set_option('graph_iters', gi)
        print(f"N={N} auto tiling graph_iters={gi}: {solve():.1f} us/iter", flush=True)
    chunk.close()
