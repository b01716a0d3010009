# A/B: CG as two kernels per iteration (graphs) vs ONE persistent cooperative kernel (option cg_persist)
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
sizes = [int(a) for a in sys.argv[1:]] or [256, 1024, 2048, 4096, 8192]
for N in sizes:
    iters = 2000 if N <= 1024 else (600 if N <= 4096 else 300)
    res = {}
    for persist in (0, 1, 0, 1):
        s = classic_settings(N, steps=1, solver='cg', maxiters=iters)
        chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
        chunk.set_option("cg_persist", persist)
        rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
        best, info = None, None
        for _ in range(3):
            chunk.copy_field("energy", "energy0")
            tl.haloupdate(chunk, s, 1, ["energy", "density"])
            info = chunk.cg_solve(s, rx, ry)
            m_ = info['solve_ms'] / info['iters']
            best = m_ if best is None else min(best, m_)
        res.setdefault(persist, []).append((best * 1e3, info['iters'], info['error']))
        chunk.close()
    a, b = min(res[0]), min(res[1])
    same = res[0][0][1:] == res[1][0][1:]
    print(f"[persist] N={N} iters={a[1]}  two-kernel {a[0]:.2f} us/iter  persistent {b[0]:.2f} us/iter  ({100 * (a[0] / b[0] - 1):+.1f} %)  "
          f"G cell-it/s {N * N / a[0] / 1e3:.2f} -> {N * N / b[0] / 1e3:.2f}  identical error/iters: {same}  all: {res}", flush=True)
