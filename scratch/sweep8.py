# tail-chunk sweep at 4096^2 (and 2048 x 4096 half tile): CG us/iter and kernel A/B times
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
for nx, ny in ((4096, 4096), (4096, 2048), (8192, 8192)):
    s = classic_settings(nx, ny=ny, steps=1, solver='cg', maxiters=400)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    def run(tag):
        best = None
        for _ in range(3):
            chunk.copy_field("energy", "energy0")
            info = chunk.cg_solve(s, rx, ry)
            m_ = info['solve_ms'] / info['iters']
            best = m_ if best is None else min(best, m_)
        ka = min(chunk.time_kernel('cg_fused_w', 30) for _ in range(2)); kb = min(chunk.time_kernel('cg_fused_r', 30) for _ in range(2))
        print(f"{nx}x{ny} {tag:40s} us/iter={best*1e3:.1f} A={ka*1e3:.1f} B={kb*1e3:.1f}", flush=True)
    run('baseline')
    for frac in (0.06, 0.12, 0.25):
        for tr, ptr in ((2, 4), (4, 4), (1, 2), (2, 8)):
            chunk.set_option('tail_frac', frac); chunk.set_option('tail_rows', tr); chunk.set_option('pw_tail_rows', ptr)
            run(f'tail_frac={frac} tail_rows={tr} pw_tail_rows={ptr}')
    chunk.set_option('tail_frac', 0.0)
    run('baseline again')
    chunk.close()
