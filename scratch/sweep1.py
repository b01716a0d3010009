import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
try:
    from cuda import cudart
    err, prop = cudart.cudaGetDeviceProperties(0)
    print('l2CacheSize', prop.l2CacheSize, 'persistingL2CacheMaxSize', prop.persistingL2CacheMaxSize,
          'accessPolicyMaxWindowSize', prop.accessPolicyMaxWindowSize, 'SMs', prop.multiProcessorCount,
          'regs/SM', prop.regsPerMultiprocessor, 'smem/SM', prop.sharedMemPerMultiprocessor)
except Exception as e:
    print('cuda-python props failed', e)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
s = classic_settings(N, steps=1, solver='cg', maxiters=400)
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
    print(f"{tag:40s} iters={info['iters']} ms/iter={best:.4f}  Gcellit/s={N*N/best/1e6:.2f} err={info['error']:.6e}", flush=True)
solve('baseline')
for pw in (2, 3, 4):
    chunk.set_option('pw_blocks_per_sm', pw); solve(f'pw_blocks_per_sm={pw}')
chunk.set_option('pw_blocks_per_sm', 4)
for g in (2, 32):
    chunk.set_option('graph_iters', g); solve(f'graph_iters={g}')
chunk.set_option('graph_iters', 8)
chunk.set_option('use_graph', 0); solve('no graph'); chunk.set_option('use_graph', 1)
FIELD = {'r': 6, 'w': 7, 'u': 3, 'kx': 8}
for fname in ('r', 'w', 'u'):
    chunk.set_option('l2_persist_field', FIELD[fname])
    for mb in (30, 50, 70, 90, 110, 130):
        chunk.set_option('l2_persist_mb', mb); solve(f'l2 persist {fname} {mb} MB')
chunk.set_option('l2_persist_field', FIELD['r'])
for scale in (0.6, 0.8, 1.2, 1.5):
    chunk.set_option('l2_hit_scale', scale)
    chunk.set_option('l2_persist_mb', 90); solve(f'l2 persist r 90MB hit_scale {scale}')
chunk.set_option('l2_persist_mb', 0)
for k in ('cg_fused_w', 'cg_fused_r', 'cheby_fused', 'ppcg_inner', 'cg_fused_w_nou'):
    print(k, chunk.time_kernel(k, 30))
