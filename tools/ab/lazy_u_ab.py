# A/B: CG kernel A advancing u every launch (cg_lazy_u = 0) vs every second launch with both pending updates (1, default)
import sys, os
sys.path.insert(0, os.getcwd())
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
from tealeaf_jl_b200.solvers import get_solver
sizes = [(int(a), int(b)) for a, b in (x.split(":") for x in (sys.argv[1] if len(sys.argv) > 1 else "4096:1500,8192:600,2048:1500,16384:200").split(","))]
for N, iters in sizes:
    s = classic_settings(N, steps=1, solver='cg', maxiters=iters)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    variants = [{"cg_lazy_u": 0}, {"cg_lazy_u": 1}, {"cg_lazy_u": 0}, {"cg_lazy_u": 1}]
    if N <= 4096:
        variants += [{"cg_lazy_u": 1, "cg_lazy_heavy_ctas": 3}, {"cg_lazy_u": 1, "ring_stages": 4}, {"cg_lazy_u": 0, "ring_stages": 4}]
    else:
        variants += [{"cg_lazy_u": 1, "ring_stages": 3}]
    for opts in variants:
        chunk.set_option("ring_stages", -1)
        chunk.set_option("cg_lazy_heavy_ctas", 2)
        for k, v in opts.items():
            chunk.set_option(k, v)
        chunk.cg_init(s.coefficient, rx, ry)
        ka = min(chunk.time_kernel('cg_fused_w', 30) for _ in range(3))
        ko = min(chunk.time_kernel('cg_fused_w_odd', 30) for _ in range(3))
        kb = min(chunk.time_kernel('cg_fused_r', 30) for _ in range(3))
        best = None
        for _ in range(2):
            chunk.copy_field("energy", "energy0")
            tl.haloupdate(chunk, s, 1, ["energy", "density"])
            info = get_solver('cg').solve(chunk, s, rx, ry)
            if best is None or info["solve_ms"] < best["solve_ms"]:
                best = info
        print(f"[lazy_u_ab] N={N} {opts} (u mode {int(chunk.get_option('cg_u_mode'))}, ring {int(chunk.get_option('ring_stages_effective'))}, {int(chunk.get_option('cg_a_blocks_per_sm'))} CTAs/SM): "
              f"kernel A {ka * 1e3:.1f} us / odd launches {ko * 1e3:.1f} us, B {kb * 1e3:.1f} us; solve {best['iters']} iterations "
              f"{best['solve_ms']:.1f} ms = {1e3 * best['solve_ms'] / best['iters']:.2f} us/iteration, error {best['error']!r}", flush=True)
    chunk.close()
