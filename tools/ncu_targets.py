#!/usr/bin/env python
"""ncu_targets.py -- launches the hot kernels a few times each through tl_time_kernel so that one
`ncu --set full -k regex:... ` pass can capture them (never a bench value: ncu serialises and replays).

  python tools/ncu_targets.py --n 4096 --kernels cheby_pair,ppcg_pair,ppcg_inner,jacobi_fused,cheby_fused
  python tools/ncu_targets.py --n 8192 --kernels cg_fused_w,cg_fused_r        (ring depth 4 beyond 8192^2)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--kernels", default="cg_fused_w,cg_fused_r")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    import tealeaf_jl_b200 as tl
    from tealeaf_jl_b200.decks import classic_settings
    from tealeaf_jl_b200.device import DeviceChunk
    s = classic_settings(args.n, ny=args.ny or args.n, steps=1, solver="cg", maxiters=200)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    for kv in args.opt:
        k, _, v = kv.partition("=")
        chunk.set_option(k, float(v))
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    chunk.cg_init(s.coefficient, rx, ry)            # kx, ky, u, p, r hold real data
    for name in args.kernels.split(","):
        ms = chunk.time_kernel(name, args.reps)
        print(f"[ncu_targets] {name} @ {args.n}x{args.ny or args.n}: {ms * 1e3:.1f} us per launch (under ncu: not a bench value)", flush=True)
    chunk.close()


if __name__ == "__main__":
    main()
