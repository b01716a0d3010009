#!/usr/bin/env python
"""Target of tools/sanitize.sh: every kernel family of libtealeaf_b200 once, on meshes small enough
for compute-sanitizer (memcheck / racecheck / synccheck slow kernels down 10-100x).  Odd sizes on
purpose: partial strips, a single last column, chunks shorter than the ring depth."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import tealeaf_jl_b200 as tl  # noqa: E402
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk  # noqa: E402

CASES = [("cg", 97, 61, {}, {}), ("cg", 130, 40, {}, {"cg_persist": 1}), ("cg", 75, 90, {}, {"b_ring": 6}),
         ("cg", 64, 33, {}, {"b_ring": 8}), ("cheby", 97, 61, {}, {}), ("ppcg", 97, 61, {"ppcginnersteps": 4}, {}),
         ("jacobi", 50, 45, {"maxiters": 120}, {}), ("cg", 1, 40, {}, {}), ("cg", 200, 3, {"halodepth": 3}, {}),
         # round 2: TMA ring of kernel A, odd PPCG inner count (pairs + trailing step), 5-slot pair ring, Chebyshev pairs
         ("cg", 130, 70, {}, {"a_tma": 4}), ("cg", 63, 40, {}, {"a_tma": 3}), ("ppcg", 97, 61, {"ppcginnersteps": 5}, {}),
         ("cheby", 130, 70, {}, {"pair_stages": 5}), ("ppcg", 70, 90, {"ppcginnersteps": 6}, {"pair_stages": 5}),
         # lazy u update of the CG loop (default on: the cases above run it at ring depth 3, two CTAs per SM): the other flavours
         ("cg", 97, 61, {}, {"cg_lazy_u": 0}), ("cg", 131, 45, {}, {"cg_lazy_heavy_ctas": 3}), ("cg", 67, 80, {}, {"ring_stages": 4}),
         ("cg", 65, 33, {"maxiters": 7}, {"ring_stages": 6})]


def main():
    for solver, nx, ny, over, opts in CASES:
        s = classic_settings(nx, ny=ny, steps=1, solver=solver, **over)
        chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)          # device-side state painter
        for k, v in opts.items():
            chunk.set_option(k, v)
        recs, final = tl.diffuse(chunk, s, geom)                        # fused solve + finalise + field summary
        u = chunk.get_field("u")
        assert np.isfinite(u).all()
        print(f"sanitize target: {solver} {nx}x{ny} {opts} iters={recs[0]['iters']} temp={final['temp']:.12g}", flush=True)
        chunk.close()
    # the per-function entry points (kernels.jl / CG.jl names) on one more odd mesh
    s = classic_settings(70, ny=50, steps=1, solver="cg")
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    recs, final = tl.diffuse(chunk, s, geom, stepwise=True)
    print(f"sanitize target: stepwise cg 70x50 iters={recs[0]['iters']}", flush=True)
    chunk.close()
    print("sanitize target OK")


if __name__ == "__main__":
    main()
