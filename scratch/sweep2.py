import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
from oracle.oracle import OracleChunk

# correctness of every ring flavour on awkward shapes
for (nx, ny) in ((130, 77), (64, 64), (257, 19)):
    so = classic_settings(nx, ny=ny, steps=1)
    oc, og = tl.initialiseapp(so, backend=OracleChunk)
    orec, _ = tl.diffuse(oc, so, og)
    for ring in (0, 3, 4, 6):
        s = classic_settings(nx, ny=ny, steps=1)
        c, g = tl.initialiseapp(s, backend=DeviceChunk)
        c.set_option('ring_stages', ring)
        rec, _ = tl.diffuse(c, s, g)
        err = np.abs(c.get_field('u') - oc.get_field('u')).max() / np.abs(oc.get_field('u')).max()
        perr = np.abs(c.get_field('p') - oc.get_field('p')).max() / np.abs(oc.get_field('u')).max()
        print(f'{nx}x{ny} ring={ring}: iters {rec[0]["iters"]} vs {orec[0]["iters"]}  u err {err:.2e} p err {perr:.2e}', flush=True)
        c.close()

for N in (4096, 16384):
    s = classic_settings(N, steps=1, solver='cg', maxiters=200 if N == 4096 else 60)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    def solve(tag):
        tl.haloupdate(chunk, s, 1, ["energy", "density"])
        chunk.copy_field("energy", "energy0")
        best = None
        for _ in range(3):
            info = chunk.cg_solve(s, rx, ry)
            ms = info['solve_ms'] / info['iters']
            best = ms if best is None else min(best, ms)
        ka = chunk.time_kernel('cg_fused_w', 20)
        print(f"N={N} {tag:28s} ms/iter={best:.4f} Gcellit/s={N*N/best/1e6:.2f}  kernelA={ka*1e3:.1f}us ({64*N*N/ka/1e6/1e3:.0f} GB/s phys)", flush=True)
    for ring in (0, 3, 4, 6):
        chunk.set_option('ring_stages', ring)
        solve(f'ring={ring}')
    chunk.set_option('ring_stages', 4)
    for bps in (1, 2):
        chunk.set_option('blocks_per_sm', bps); solve(f'ring=4 blocks_per_sm={bps}')
    chunk.set_option('ring_stages', 3)
    for cr in (8, 16, 32, 64):
        chunk.set_option('chunk_rows', cr); solve(f'ring=3 chunk_rows={cr}')
    chunk.close()
