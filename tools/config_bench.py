#!/usr/bin/env python
"""config_bench.py -- timing of the BASELINE.json configurations that are not bench.py's headline
line (configs[2..4]: Chebyshev 4096^2, PPCG 8192^2, CG weak scaling), at 1..8 GPUs, with either
multi-GPU data path (comm = fused | nccl).  Launch with torch.distributed.run for N > 1:

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/config_bench.py \
      --solver cheby --global 4096 --max-iters 2000 [--comm nccl] [--inner 10] [--tile 16384]
      [--halo-depth 4 --ppcg-halo-depth 4]      (PPCG: one tile exchange per 4 inner steps)

Prints one JSON line per (case, comm) on rank 0: global cell-iterations/s from the device time of
the solve (CUDA events on the solve stream, max over ranks), per-kernel-launch average, and the
algorithmic bytes per cell-iteration figure of SURVEY.md §8(d)."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ALG_BYTES = {"cg": 104, "cheby": 88}   # per cell-iteration; PPCG: 128 outer + 80 per inner step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--solver", default="cg", choices=["cg", "cheby", "ppcg"])
    ap.add_argument("--global", dest="glob", type=int, default=0, help="global mesh side (strong scaling)")
    ap.add_argument("--tile", type=int, default=0, help="cells per GPU per side (weak scaling)")
    ap.add_argument("--max-iters", type=int, default=1000)
    ap.add_argument("--inner", type=int, default=10)
    ap.add_argument("--ppcg-halo-depth", default="0",
                    help="PPCG: tile exchange every k inner steps (0 = halo_depth; 1 = every step); a comma list "
                         "times every depth in one process")
    ap.add_argument("--halo-depth", type=int, default=2)
    ap.add_argument("--comm", default="fused,nccl")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--opt", action="append", default=[], help="tl_set_option name=value (repeatable)")
    ap.add_argument("--prof", action="store_true",
                    help="kernel-boundary micro-profile (globaltimer stamps in the kernel tails): averages per kernel in the JSON line")
    args = ap.parse_args()

    import torch
    import tealeaf_jl_b200 as tl
    from tealeaf_jl_b200 import dist as tld
    from tealeaf_jl_b200.chunk import HostGeometry
    from tealeaf_jl_b200.device import DeviceChunk
    from tealeaf_jl_b200.solvers import get_solver
    from tealeaf_jl_b200.decks import classic_settings

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    px, py = tld.grid_for(world)
    if args.tile:
        nx, ny = args.tile * px, args.tile * py
    else:
        nx = ny = args.glob or 4096
    over = {"maxiters": args.max_iters, "halodepth": args.halo_depth}
    depths = [int(k) for k in str(args.ppcg_halo_depth).split(",")] if args.solver == "ppcg" else [0]
    comms = args.comm.split(",") if world > 1 else ["single"]
    for comm, depth in [(cm, k) for cm in comms for k in depths]:
        if args.solver == "ppcg":
            over["ppcginnersteps"] = args.inner
            over["ppcghalodepth"] = depth
        s = classic_settings(nx, ny=ny, steps=1, solver=args.solver, **over)
        if world > 1:
            chunk, geom, _ = tld.create_tile(s, dist, local_rank, options={"comm_fused": 1 if comm == "fused" else 0})
        else:
            chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
        for kv in args.opt:
            k, _, v = kv.partition("=")
            chunk.set_option(k, float(v))
        rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
        solver = get_solver(s.solver)
        best = None
        for _ in range(args.reps + 1):           # first pass builds the graphs
            chunk.copy_field("energy", "energy0")
            tl.haloupdate(chunk, s, 1, ["energy", "density"])
            info = solver.solve(chunk, s, rx, ry)
            if best is None or info["solve_ms"] < best["solve_ms"]:
                best = info
        prof = None
        if args.prof:      # one more solve with the stamps on (they cost a few globaltimer reads per kernel)
            chunk.set_option("prof", 1)
            chunk.copy_field("energy", "energy0")
            tl.haloupdate(chunk, s, 1, ["energy", "density"])
            pinfo = solver.solve(chunk, s, rx, ry)
            prof = {k: chunk.get_option("prof_" + k) for k in ("kernels", "body_us", "sum_us", "xchg_us", "gap_us", "fence_us")}
            prof["solve_ms_with_stamps"] = pinfo["solve_ms"]
            chunk.set_option("prof", 0)
            if dist is not None:
                allp = [None] * world
                dist.all_gather_object(allp, prof)
                prof = {"rank0": allp[0], "max_over_ranks": {k: max(p[k] for p in allp) for k in prof},
                        "min_over_ranks": {k: min(p[k] for p in allp) for k in prof}}
        ms = best["solve_ms"]
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        cells = nx * ny
        if args.solver == "ppcg":
            outer = best["cheby_iters"]
            work_iters = best["cg_iters"] + outer + best["inner_total"]
            alg = (best["cg_iters"] * 104 + outer * 128 + best["inner_total"] * 80) * cells
        else:
            work_iters = best["iters"]
            alg = (best["cg_iters"] * 104 + best["cheby_iters"] * 88) * cells
        if rank == 0:
            print(json.dumps({
                "solver": args.solver, "global_cells": [nx, ny], "n_gpus": world, "decomposition": f"{px}x{py}",
                "comm": comm, "options": args.opt, "iters": best["iters"], "cg_iters": best["cg_iters"], "cheby_or_outer_iters": best["cheby_iters"],
                "inner_steps_total": best["inner_total"], "halo_depth_k": best.get("halo_depth_k"), "error": best["error"], "solve_ms": ms,
                "cell_iterations_per_s": cells * work_iters / (ms * 1e-3),
                "us_per_sweep": 1e3 * ms / max(work_iters, 1),
                "algorithmic_gbs_per_gpu": alg / (ms * 1e-3) / 1e9 / world,
                "kernel_launches": best["kernel_launches"],
                **({"boundary_profile_us_per_kernel": prof} if prof else {}),
            }), flush=True)
        chunk.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
