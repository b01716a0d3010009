import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
N = 4096
for solver in ('cheby', 'ppcg'):
    for ring in (0, -1):
        s = classic_settings(N, steps=1, solver=solver, maxiters=400)
        chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
        chunk.set_option('ring_stages', ring)
        if ring == 0:
            chunk.set_option('chunk_rows', 0); chunk.set_option('pw_chunk_rows', 0)
        recs, final = tl.diffuse(chunk, s, geom)
        r = recs[0]
        k = 'cheby_fused' if solver == 'cheby' else 'ppcg_inner'
        kt = chunk.time_kernel(k, 20)
        phys = 56 if solver == 'cheby' else 64
        print(f"{solver} ring={ring}: iters={r['iters']} cg={r['cg_iters']} cheby/outer={r['cheby_iters']} inner={r['inner_total']} solve_ms={r['solve_ms']:.2f} err={r['error']:.3e}  kernel {k}={kt*1e3:.1f}us ({phys*N*N/kt/1e9:.2f} TB/s phys)", flush=True)
        chunk.close()
