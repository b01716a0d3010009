# mid-size tiles (what a 4096^2 / 8192^2 mesh becomes on 2..8 GPUs): wave quantisation of the chunk length
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
shapes = [(2048, 1024), (2048, 2048), (4096, 2048), (1024, 1024), (1024, 512)]
for nx, ny in shapes:
    s = classic_settings(nx, ny=ny, steps=1, solver='cg', maxiters=300)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    for cr in (-1, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 20):
        chunk.set_option("chunk_rows", cr)
        ka = min(chunk.time_kernel('cg_fused_w', 40) for _ in range(3))
        kc = min(chunk.time_kernel('cheby_fused', 40) for _ in range(3))
        ki = min(chunk.time_kernel('ppcg_inner', 40) for _ in range(3))
        print(f"[tile] {nx}x{ny} chunk_rows={cr:3d}  A={ka * 1e3:6.2f} us  cheby={kc * 1e3:6.2f} us  inner={ki * 1e3:6.2f} us", flush=True)
    chunk.set_option("chunk_rows", -1)
    for pcr in (-1, 4, 8, 12, 16, 20, 24, 32):
        chunk.set_option("pw_chunk_rows", pcr)
        kb = min(chunk.time_kernel('cg_fused_r', 40) for _ in range(3))
        print(f"[tile] {nx}x{ny} pw_chunk_rows={pcr:3d}  B={kb * 1e3:6.2f} us", flush=True)
    chunk.close()
