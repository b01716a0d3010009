"""CLI -- mirror of the reference's `run.jl` (same flags, `run.jl:5-23`).

    python -m tealeaf_jl_b200.run -i decks/tea_bm_small.in [-s cg|cheby|ppcg] [-x N] [-y N] [--stepwise]
"""
from __future__ import annotations

import argparse
import json
import logging

from .app import diffuse, initialiseapp
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
    chunk, geom = initialiseapp(settings, device=args.device)              # run.jl:45
    records, final = diffuse(chunk, settings, geom, stepwise=args.stepwise)  # run.jl:47
    for r in records:
        print(json.dumps({k: v for k, v in r.items() if k != "summary"}))
    print(json.dumps({"final_summary": final}))


if __name__ == "__main__":
    main()
