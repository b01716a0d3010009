"""Multi-GPU parity check, run under torchrun (one rank per GPU):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_check.py
Every solver is run tiled over the N GPUs and compared on rank 0 with the single-chunk CPU
oracle on the same deck.  Exits non-zero on any mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import tealeaf_jl_b200 as tl  # noqa: E402
from tealeaf_jl_b200 import dist as tld  # noqa: E402
from tealeaf_jl_b200.decks import classic_settings  # noqa: E402


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    # (solver, nx, ny, steps, settings overrides, tl_set_option values)
    F, N = {"comm_fused": 1}, {"comm_fused": 0}
    cases = [("cg", 256, 192, 2, {}, F), ("cg", 130, 77, 1, {}, F), ("cheby", 192, 256, 1, {}, F),
             ("ppcg", 192, 160, 1, {"ppcginnersteps": 6}, F), ("ppcg", 131, 150, 1, {"ppcginnersteps": 5}, F),
             ("cheby", 129, 67, 1, {}, F), ("cg", 512, 512, 1, {"maxiters": 300}, F),
             ("jacobi", 160, 130, 1, {"maxiters": 120}, F),
             # the schedules the defaults replaced: one kernel + one exchange per Chebyshev iteration / PPCG inner step,
             # and the matrix-powers groups (exchange every halo_depth inner steps)
             ("cheby", 192, 256, 1, {}, {**F, "pair_tiled": 0}), ("ppcg", 192, 160, 1, {"ppcginnersteps": 6, "ppcghalodepth": 1}, F),
             ("ppcg", 131, 150, 1, {"ppcginnersteps": 5, "ppcghalodepth": 2}, F),
             # larger tiles: several chunks per strip, both pair kernels
             ("cheby", 1024, 768, 1, {"maxiters": 600}, F), ("ppcg", 1024, 768, 1, {"maxiters": 500}, F),
             # the older halo-pull + NCCL path (comm_fused = 0) stays available for A/B measurements
             ("cg", 256, 192, 1, {}, N), ("cheby", 192, 256, 1, {}, N), ("ppcg", 192, 160, 1, {"ppcginnersteps": 6}, N)]
    if len(sys.argv) > 1:
        cases = [c for c in cases if c[0] in sys.argv[1:]]
    failures = 0
    for solver, nx, ny, steps, over, opts in cases:
        fused = opts["comm_fused"]
        s = classic_settings(nx, ny=ny, steps=steps, solver=solver, **over)
        chunk, geom, (px, py) = tld.create_tile(s, dist, local_rank, options=opts)
        summaries = []
        recs, final = tl.diffuse(chunk, s, geom,
                                 on_step=lambda rec: summaries.append(chunk.fieldsummary(geom.cell_volume)))
        u = tld.gather_field(chunk, "u", s, dist)
        e = tld.gather_field(chunk, "energy", s, dist)
        chunk.close()
        if rank == 0:
            from oracle.oracle import OracleChunk
            so = classic_settings(nx, ny=ny, steps=steps, solver=solver, **over)
            threads = 8 if nx * ny > 500000 else 1
            oc, og = tl.initialiseapp(so, backend=lambda *a, **k: OracleChunk(*a, threads=threads, **k))
            osum = []
            orecs, ofinal = tl.diffuse(oc, so, og, on_step=lambda rec: osum.append(oc.fieldsummary(og.cell_volume)))
            uo, eo = oc.get_field("u"), oc.get_field("energy")
            hd = s.halodepth
            err_u = np.abs(u[hd:-hd, hd:-hd] - uo[hd:-hd, hd:-hd]).max() / np.abs(uo).max()
            err_e = np.abs(e[hd:-hd, hd:-hd] - eo[hd:-hd, hd:-hd]).max() / np.abs(eo).max()
            its = [r["iters"] for r in recs]
            oits = [r["iters"] for r in orecs]
            serr = max(abs(a / b - 1) for sa, sb in zip(summaries, osum) for a, b in zip(sa, sb))
            slack = 1 if solver == "cg" else 0
            ok = all(abs(a - b) <= slack for a, b in zip(its, oits)) and err_u < 1e-9 and err_e < 1e-9 and serr < 1e-10
            print(f"[mgpu {world} GPUs {px}x{py} {'fused' if fused else 'nccl '}] {solver} {nx}x{ny} {over} {({k: v for k, v in opts.items() if k != 'comm_fused'})}: iters {its} oracle {oits}  "
                  f"u err {err_u:.2e}  energy err {err_e:.2e}  summary err {serr:.2e}  {'OK' if ok else 'FAIL'}", flush=True)
            failures += 0 if ok else 1
    flag = torch.tensor([failures], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
