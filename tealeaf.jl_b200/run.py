"""CLI -- mirror of the reference's `run.jl` (same flags, `run.jl:5-23`).

    python -m tealeaf_jl_b200.run -i decks/tea_bm_small.in [-s cg|cheby|ppcg|jacobi] [-x N] [-y N] [-O dump]
                                  [--gpus N] [--tea-out tea.out] [--stepwise]
"""
from __future__ import annotations

import argparse
import json
import logging
import time

from .app import diffuse, initialiseapp, write_tea_out
from .settings import parse_settings


def main(argv=None):
    ap = argparse.ArgumentParser(description="TeaLeaf heat conduction on B200 (libtealeaf_b200)")
    ap.add_argument("--solver", "-s", choices=["cg", "cheby", "ppcg", "jacobi"], type=str.lower,
                    help="Can be 'cg', 'cheby', 'ppcg', or 'jacobi'")      # run.jl:6-9
    ap.add_argument("-x", type=int, help="Number of x cells")              # run.jl:10-12
    ap.add_argument("-y", type=int, help="Number of y cells")              # run.jl:13-15
    ap.add_argument("-i", "--in-file", default="tea.in", help="Settings input file")   # run.jl:16-19
    ap.add_argument("-O", "--debug-out", help="File to print debug state to")  # run.jl:20-22
    ap.add_argument("--stepwise", action="store_true", help="drive the solve kernel by kernel (per-function ABI)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1, help="spread the one Chunk over N GPUs of this process (tl_create_multi)")
    ap.add_argument("--tea-out", help="write an upstream-style tea.out report (steps, iterations, field summaries, QA)")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    settings = parse_settings(args.in_file)                                # run.jl:26
    if args.solver:
        settings.solver = args.solver                                      # run.jl:27-34
    if args.x:
        settings.xcells = args.x
    if args.y:
        settings.ycells = args.y
    if args.debug_out:
        settings.debugfile = args.debug_out                                # run.jl:41-43
    settings.recompute_spacing()                                           # Appendix A #23
    t0 = time.perf_counter()
    if args.gpus > 1:
        from .device import DeviceChunk
        chunk, geom = initialiseapp(settings, backend=DeviceChunk.multi, ngpus=args.gpus)
    else:
        chunk, geom = initialiseapp(settings, device=args.device)          # run.jl:45
    records, final = diffuse(chunk, settings, geom, stepwise=args.stepwise)  # run.jl:47
    if args.tea_out:
        write_tea_out(args.tea_out, settings, records, final, time.perf_counter() - t0)
    for r in records:
        print(json.dumps({k: v for k, v in r.items() if k != "summary"}))
    print(json.dumps({"final_summary": final}))


if __name__ == "__main__":
    main()
