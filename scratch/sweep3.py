import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
for N in (4096, 16384):
    s = classic_settings(N, steps=1, solver='cg', maxiters=200 if N == 4096 else 40)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    def solve(tag, full=True):
        ms = float('nan')
        if full:
            tl.haloupdate(chunk, s, 1, ["energy", "density"])
            chunk.copy_field("energy", "energy0")
            best = None
            for _ in range(2):
                info = chunk.cg_solve(s, rx, ry)
                m_ = info['solve_ms'] / info['iters']
                best = m_ if best is None else min(best, m_)
            ms = best
        ka = chunk.time_kernel('cg_fused_w', 20)
        kb = chunk.time_kernel('cg_fused_r', 20)
        print(f"N={N} {tag:34s} ms/iter={ms:.4f} Gcellit/s={N*N/ms/1e6:.2f}  A={ka*1e3:.1f}us ({64*N*N/ka/1e9:.2f} TB/s)  B={kb*1e3:.1f}us ({24*N*N/kb/1e9:.2f} TB/s)", flush=True)
    for ring in (0, 3, 4):
        chunk.set_option('ring_stages', ring)
        for cr in ((2, 4, 8, 12, 16, 24) if N == 4096 else (4, 8, 16)):
            chunk.set_option('chunk_rows', cr); solve(f'ring={ring} chunk_rows={cr}', full=False)
    chunk.set_option('ring_stages', 3); chunk.set_option('chunk_rows', 8)
    for pcr in (0, 4, 8, 16, 32, 64):
        chunk.set_option('pw_chunk_rows', pcr); solve(f'ring=3 cr=8 pw_chunk_rows={pcr}', full=False)
    for pbs in (2, 3, 4):
        chunk.set_option('pw_blocks_per_sm', pbs); chunk.set_option('pw_chunk_rows', 0); solve(f'pw_blocks_per_sm={pbs} one wave', full=False)
    chunk.set_option('pw_blocks_per_sm', 4); chunk.set_option('pw_chunk_rows', 8)
    solve('ring=3 cr=8 pw_cr=8 FULL')
    chunk.set_option('graph_iters', 32); solve('... graph_iters=32 FULL')
    chunk.close()
