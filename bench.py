#!/usr/bin/env python
"""bench.py -- solver throughput of the B200-native TeaLeaf path (BASELINE.json metric:
solver cell-iterations/s and HBM GB/s vs roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N = 1 (default): BASELINE.json configs[1] -- the tea_bm_5-style 4096 x 4096 mesh, CG solver,
one B200; a "step" is one timestep = one CG solve to the reference's convergence test
(about 4.7k iterations).  K = 10 steps is the benchmark's 10 timesteps.
N > 1 (under torchrun): BASELINE.json configs[4] -- weak scaling, 16384 x 16384 cells per GPU,
CG with the iteration count capped per step, 2-D domain decomposition.

Prints ONE JSON line (rank 0).  `value` = cells x iterations / device time with all fields
resident in HBM; `e2e` = the same through the C-ABI with HOST buffers, uploads of that step's
inputs and the download of its result inside the timed region.  `--impl reference` times the
CPU oracle (the reference is Julia and cannot run here) with all host threads
on a bounded sample of the same workload (and never loads the CUDA library).  Every line also carries
`other_configs`: BASELINE.json configs[2] (Chebyshev, 4096 x 4096 global) and configs[3] (PPCG, 8192 x 8192
global, tile exchange every step vs every `halo_depth` steps) solved once on the N GPUs (strong scaling), and at
N = 1 `weak_tile_16384`: the 16384 x 16384 tile of the N > 1 workload solved alone, the like-for-like base of
the weak-scaling efficiency.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "solver cell-iterations/sec"
UNIT = "cell-iterations/s"
CG_ALG_BYTES = 104          # w! 32 + ur! 48 + p! 24 bytes per cell-iteration
KERNEL_A_ALG_BYTES = 80     # k_cg_fused_w covers w! (32) + p! (24) + the u half of ur! (24)
KERNEL_A_PHYS_BYTES = 64    # what it physically moves: read r,p,u,kx,ky; write p,u,w
KERNEL_B_ALG_BYTES = 24     # k_cg_fused_r: the r half of ur!


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def classic(nx, ny, steps, maxiters=10000, solver="cg"):
    from tealeaf_jl_b200.decks import classic_settings
    return classic_settings(nx, ny=ny, steps=steps, solver=solver, maxiters=maxiters)


# ------------------------------------------------------------------------------------------
# CPU arm (oracle): used for cpu_baseline and for --impl reference
# ------------------------------------------------------------------------------------------
def cpu_cg_sample(n, iters, threads, reps=1):
    """`iters` CG iterations of the oracle at n x n (classic deck) on `threads` host threads.
    Returns (cell-iterations/s, seconds per rep)."""
    import tealeaf_jl_b200 as tl
    from oracle.oracle import OracleChunk
    s = classic(n, n, 1)
    chunk, geom = tl.initialiseapp(s, backend=lambda *a, **k: OracleChunk(*a, threads=threads, **k))
    tl.haloupdate(chunk, s, 1, ["energy", "density"])
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    best = None
    for _ in range(reps):
        rro = chunk.cg_init(s.coefficient, rx, ry)
        chunk.haloupdate(["u", "p"], 1)
        chunk.copyu()
        t0 = time.perf_counter()
        chunk.cg_fixed_iters(rro, iters)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    chunk.close()
    return n * n * iters / best, best


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle port; the Julia reference
    cannot execute in this environment) on all host threads, bounded sample per step.  Never touches
    the CUDA library.  The metric is per cell-iteration, so the sample is the same at every N: CG
    iterations on the 4096 x 4096 classic deck (named in config.sample -- at N > 1 the B200 arm's
    workload is 16384^2 cells per GPU, which the CPU would need 24 GB and minutes per step for)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import load
    threads = load().tlo_max_threads()
    n, iters = 4096, 20
    import tealeaf_jl_b200 as tl
    from oracle.oracle import OracleChunk
    s = classic(n, n, 1)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2

    def make(nthreads):
        chunk, geom = tl.initialiseapp(s, backend=lambda *a, **k: OracleChunk(*a, threads=nthreads, **k))
        tl.haloupdate(chunk, s, 1, ["energy", "density"])

        def step(niters):
            rro = chunk.cg_init(s.coefficient, rx, ry)
            chunk.haloupdate(["u", "p"], 1)
            chunk.copyu()
            t0 = time.perf_counter()
            chunk.cg_fixed_iters(rro, niters)
            return time.perf_counter() - t0
        return chunk, step

    chunk, step = make(threads)
    for _ in range(args.warmup):
        step(iters)
    total = sum(step(iters) for _ in range(args.steps))
    chunk.close()
    value = n * n * iters * args.steps / total
    # the reference itself is serial (no threading anywhere, CG.jl:82-90): one thread, a shorter sample
    chunk1, step1 = make(1)
    step1(2)
    serial_iters = 6
    t1 = step1(serial_iters)
    chunk1.close()
    serial = n * n * serial_iters / t1
    sample = f"{iters} CG iterations per step of the {n}x{n} classic deck (w!, ur!, p!, halo), OpenMP oracle, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "solver": "cg", "sample": sample,
                   "note": "CPU oracle port of the reference algorithm (the Julia reference is not runnable here); the metric is "
                           "per cell-iteration, so the bounded CPU sample is the 4096x4096 deck at every N"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "cpu_baseline_serial": {"value": serial, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"{serial_iters} CG iterations of the same deck on one thread ({t1:.1f} s): what the "
                                          "reference is (no threading anywhere)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_name(ngpus):
    if ngpus == 1:
        return "tea_bm_5-style 4096x4096 classic deck, CG to convergence (eps=1e-15), one timestep per step"
    return "weak scaling: 16384x16384 cells per GPU, classic deck on the global mesh, CG capped at 200 iterations per step"


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import tealeaf_jl_b200 as tl
    from tealeaf_jl_b200.chunk import HostGeometry
    from tealeaf_jl_b200.device import DeviceChunk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    from tealeaf_jl_b200.dist import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local_rank)      # before any pinned allocation (first touch)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world == 1:
        nx = ny = 4096
        tile_nx, tile_ny, px, py, maxiters = nx, ny, 1, 1, 10000
        x0 = y0 = 0
    else:
        from tealeaf_jl_b200.dist import grid_for, tile_of
        px, py = grid_for(world)
        tile_nx = tile_ny = args.tile
        nx, ny = tile_nx * px, tile_ny * py
        x0, y0, _, _ = tile_of(rank, px, py, nx, ny)
        maxiters = args.cap_iters
    s = classic(nx, ny, args.steps, maxiters=maxiters)
    geom = HostGeometry(s, tile=(x0, y0, tile_nx, tile_ny))
    x, y = geom.x, geom.y

    # N>1: the same tile solved alone on this GPU (no neighbours, no exchange), so that the
    # weak-scaling efficiency of THIS workload can be read from one JSON line
    solo = None
    if world > 1:
        ssolo = classic(tile_nx, tile_ny, 1, maxiters=60)
        csolo = DeviceChunk(tile_nx, tile_ny, ssolo.halodepth, ssolo.maxiters, device=local_rank)
        csolo.paint_states(s, geom)
        csolo.haloupdate(["density", "energy0", "energy"], 1)
        csolo.copy_field("energy", "energy0")
        rxs, rys = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
        csolo.cg_solve(ssolo, rxs, rys)
        csolo.timer_start()
        isolo = csolo.cg_solve(ssolo, rxs, rys)
        solo_ms = csolo.timer_stop()
        solo = tile_nx * tile_ny * isolo["iters"] / (solo_ms * 1e-3)
        csolo.close()
    chunk = DeviceChunk(tile_nx, tile_ny, s.halodepth, s.maxiters, device=local_rank, rank=rank, px=px, py=py)
    if world > 1:
        from tealeaf_jl_b200.dist import connect
        connect(chunk, dist)
    rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
    # The initial state is painted on the device (tl_paint_states = setchunkstate!); the pinned
    # host copies the e2e leg uploads every step are read back from it once (Fortran (x,y) == C (y,x)).
    chunk.paint_states(s, geom)
    h_density = torch.empty((y, x), dtype=torch.float64).pin_memory()
    h_energy = torch.empty((y, x), dtype=torch.float64).pin_memory()
    chunk.get_field_raw("density", h_density.data_ptr(), x)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def upload_initial():
        chunk.paint_states(s, geom)
        chunk.haloupdate(["density", "energy0", "energy"], 1)
        chunk.copy_field("energy", "energy0")

    def timestep():
        chunk.haloupdate(["energy", "density"], 1)
        info = chunk.cg_solve(s, rx, ry)
        chunk.solvefinished(True)   # reference default check_result=true: residual! runs
        return info

    # ---- warm-up (W >= 3 timesteps; also builds the CUDA graphs) ----
    upload_initial()
    for _ in range(max(args.warmup, 0)):
        timestep()

    # ---- timed region 1: K timesteps, everything resident in HBM ----
    upload_initial()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = chunk.launch_count()
    barrier()
    chunk.timer_start()
    t0 = time.perf_counter()
    infos = [timestep() for _ in range(args.steps)]
    dev_ms = chunk.timer_stop()
    barrier()
    wall = time.perf_counter() - t0
    launches = chunk.launch_count() - l0
    clocks = sampler.stop()
    iters = sum(i["iters"] for i in infos)
    solve_ms = sum(i["solve_ms"] for i in infos)
    final_summary = chunk.fieldsummary(geom.cell_volume)

    # ---- timed region 2 (e2e): host buffers in, host result out, every step ----
    upload_initial()
    chunk.get_field_raw("energy", h_energy.data_ptr(), x)     # host copy of the evolving state
    barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        chunk.set_field_raw("density", h_density.data_ptr(), x)
        chunk.set_field_raw("energy", h_energy.data_ptr(), x)
        info = timestep()
        e2e_iters += info["iters"]
        chunk.get_field_raw("energy", h_energy.data_ptr(), x)
        chunk.fieldsummary(geom.cell_volume)
    barrier()
    e2e_wall = time.perf_counter() - t0
    field_bytes = x * y * 8

    # max over ranks of the timed durations
    def maxr(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dev_ms, wall, e2e_wall = maxr(dev_ms), maxr(wall), maxr(e2e_wall)
    cells_global = nx * ny
    value = cells_global * iters / (dev_ms * 1e-3)
    e2e_value = cells_global * e2e_iters / e2e_wall

    # ---- dominant kernel in isolation (state is scratch from here on) ----
    peak, peak_src = measured_peaks()
    tile_cells = tile_nx * tile_ny
    ka_ms = chunk.time_kernel("cg_fused_w", 30)
    kb_ms = chunk.time_kernel("cg_fused_r", 30)
    it_ms = solve_ms / max(iters, 1)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(f"k_cg_fused_w@{tile_nx}x{tile_ny}")
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": "k_cg_fused_w_ring<true, S, MINB> (CG kernel A)",
        "achieved": KERNEL_A_ALG_BYTES * tile_cells / (ka_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": KERNEL_A_ALG_BYTES * tile_cells / (ka_ms * 1e-3) / 1e9 / peak, "traffic": traffic,
        "peak_source": peak_src, "avg_launch_ms": ka_ms,
        "algorithmic_bytes_per_cell": KERNEL_A_ALG_BYTES, "physical_bytes_per_cell": KERNEL_A_PHYS_BYTES,
        "physical_gbs": KERNEL_A_PHYS_BYTES * tile_cells / (ka_ms * 1e-3) / 1e9,
        "other_kernels": {"k_cg_fused_r": {"avg_launch_ms": kb_ms,
                                           "achieved": KERNEL_B_ALG_BYTES * tile_cells / (kb_ms * 1e-3) / 1e9}},
        "iteration": {"ms": it_ms, "algorithmic_bytes_per_cell": CG_ALG_BYTES,
                      "achieved": CG_ALG_BYTES * tile_cells / (it_ms * 1e-3) / 1e9,
                      "frac_of_measured": CG_ALG_BYTES * tile_cells / (it_ms * 1e-3) / 1e9 / peak,
                      "frac_of_nominal_8TBs": CG_ALG_BYTES * tile_cells / (it_ms * 1e-3) / 1e9 / 8000.0},
    }

    # ---- the reduction-free iterations of configs[2] / configs[3] on the same tile (kernels timed alone) ----
    other = None
    if world == 1:
        try:
            def per_it(name, its_per_launch):
                ms = min(chunk.time_kernel(name, 30) for _ in range(2))
                return 1e3 * ms / its_per_launch
            c1, c2 = per_it("cheby_fused", 1), per_it("cheby_pair", 2)
            p1, p2 = per_it("ppcg_inner", 1), per_it("ppcg_pair", 2)
            other = {
                "note": "kernels of BASELINE.json configs[2]/[3] timed alone on this 4096x4096 tile (CUDA events, 30 launches); "
                        "two_per_pass = temporal blocking (k_cheby_pair_ring / k_ppcg_pair_ring), the default on a single tile",
                "chebyshev_iteration": {"one_per_pass_us": c1, "two_per_pass_us": c2, "algorithmic_bytes_per_cell": 88,
                                        "cell_iterations_per_s": tile_cells / (c2 * 1e-6),
                                        "algorithmic_gbs": 88 * tile_cells / (c2 * 1e-6) / 1e9},
                "ppcg_inner_step": {"one_per_pass_us": p1, "two_per_pass_us": p2, "algorithmic_bytes_per_cell": 80,
                                    "cell_steps_per_s": tile_cells / (p2 * 1e-6),
                                    "algorithmic_gbs": 80 * tile_cells / (p2 * 1e-6) / 1e9},
            }
        except Exception as e:      # never let the side measurement break the headline line
            other = {"error": repr(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.oracle import load
        threads = load().tlo_max_threads()
        v, secs = cpu_cg_sample(4096, 600, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"600 CG iterations of the same 4096x4096 deck ({secs:.1f} s), OpenMP oracle"}
    elif rank == 0:
        cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "N>1: measured at N=1 only"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(world), "solver": "cg", "global_cells": [nx, ny],
                       "tile_cells": [tile_nx, tile_ny], "decomposition": f"{px}x{py}",
                       "multi_gpu": ("halo cells pushed into the neighbours' memory and dot products summed through "
                                     "peer-mapped mailboxes inside the solver kernels (NVLink); no NCCL in the loops"
                                     if world > 1 else "n/a"),
                       "iterations_per_step": [i["iters"] for i in infos],
                       "l2": "working set (7 fields x %.0f MB) larger than the 126 MB L2: no flush needed" % (field_bytes / 1e6),
                       "final_summary": dict(zip(("vol", "mass", "ie", "temp"), final_summary))},
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "solve_only_ms_per_step": solve_ms / args.steps,
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * field_bytes,
                    "d2h_bytes_per_step": field_bytes + 32, "ms_per_step": 1e3 * e2e_wall / args.steps,
                    "host_numa_binding": numa},
            "gpu_launches": int(launches), "clocks": clocks,
            **({"other_configs": other} if other else {}),
            **({"same_tile_single_gpu": {"value": solo, "unit": UNIT, "note":
                "this rank's tile solved alone (1x1, 60 CG iterations incl. init) in the same job; "
                "weak-scaling efficiency of the N-GPU workload = value / (N x this)"}} if solo else {}),
        }))
    chunk.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tile", type=int, default=16384, help="N>1: cells per GPU per side")
    ap.add_argument("--cap-iters", type=int, default=200, help="N>1: CG iterations per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: NCCL's version / debug banner goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import __graft_entry__
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        __graft_entry__.build()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
