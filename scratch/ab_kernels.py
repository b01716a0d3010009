# per-kernel and per-iteration timing of the CG path at N^2 for whichever library TEALEAF_B200_LIB selects
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
tag = sys.argv[1]
for N in (4096, 8192):
    s = classic_settings(N, steps=1, solver='cg', maxiters=400)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    for kv in os.environ.get("TL_OPTS", "").split(","):
        if kv:
            chunk.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    best = None
    for _ in range(3):
        chunk.copy_field("energy", "energy0")
        info = chunk.cg_solve(s, rx, ry)
        m_ = info['solve_ms'] / info['iters']
        best = m_ if best is None else min(best, m_)
    ka = min(chunk.time_kernel('cg_fused_w', 30) for _ in range(3))
    kb = min(chunk.time_kernel('cg_fused_r', 30) for _ in range(3))
    extra = ""
    try:
        kc = min(chunk.time_kernel('cheby_fused', 30) for _ in range(2)); ki = min(chunk.time_kernel('ppcg_inner', 30) for _ in range(2))
        extra = f" cheby={kc*1e3:.1f}us ({56*N*N/kc/1e9:.2f} TB/s) inner={ki*1e3:.1f}us ({64*N*N/ki/1e9:.2f} TB/s)"
    except Exception as e:
        extra = f" ({e})"
    print(f"[{tag}] N={N} us/iter={best*1e3:.1f} A={ka*1e3:.1f}us ({64*N*N/ka/1e9:.2f} TB/s) B={kb*1e3:.1f}us ({24*N*N/kb/1e9:.2f} TB/s){extra}", flush=True)
    chunk.close()
