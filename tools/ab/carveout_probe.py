# Probe: does the L1 / shared-memory split matter for kernel A?  Run once plain and once with TEALEAF_B200_CARVEOUT=1
# (every loop kernel asks for the all-shared split: 28 KB of L1 instead of 60 KB at three CTAs of 63 KB).
import sys, os
sys.path.insert(0, os.getcwd())
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
s = classic_settings(N, steps=1, solver='cg', maxiters=100)
chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
for opts in ({"cg_lazy_u": 0}, {"cg_lazy_u": 1}):
    for k, v in opts.items():
        chunk.set_option(k, v)
    chunk.cg_init(s.coefficient, rx, ry)
    t = {k: min(chunk.time_kernel(k, 30) for _ in range(3)) * 1e3 for k in ("cg_fused_w", "cg_fused_w_odd", "cg_fused_w_nou", "cg_fused_r")}
    print(f"[carveout_probe] N={N} TEALEAF_B200_CARVEOUT={os.environ.get('TEALEAF_B200_CARVEOUT', '0')} {opts}: " +
          ", ".join(f"{k} {v:.1f} us" for k, v in t.items()), flush=True)
chunk.close()
