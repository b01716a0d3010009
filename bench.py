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

JSON_OUT = sys.stdout      # main() replaces it by a private duplicate of the original stdout
METRIC = "solver cell-iterations/sec"
UNIT = "cell-iterations/s"
CG_ALG_BYTES = 104          # w! 32 + ur! 48 + p! 24 bytes per cell-iteration
KERNEL_A_ALG_BYTES = 80     # k_cg_fused_w covers w! (32) + p! (24) + the u half of ur! (24)
KERNEL_A_PHYS_BYTES = 64    # what it physically moves: read r,p,u,kx,ky; write p,u,w
KERNEL_A_PHYS_BYTES_LAZY = 60   # u advanced every second launch with both pending updates: 72 / 48 B, 60 on average
KERNEL_B_ALG_BYTES = 24     # k_cg_fused_r: the r half of ur!


def _load_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


TRAFFIC = _load_traffic()     # ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed kernel@NXxNY


def _emit(line):
    JSON_OUT.write(line + "\n")
    JSON_OUT.flush()


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
    _emit(json.dumps({
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
# BASELINE.json configs[2] / configs[3] on the N GPUs (strong scaling) and the weak-scaling base tile
# ------------------------------------------------------------------------------------------
def other_config_legs(world, rank, local_rank, dist, peak, maxr, budget_s=75.0):
    """One bounded solve per configuration, timed on the device (CUDA events on the solve stream, max over
    ranks); the CG presteps and the Chebyshev / PPCG phase are timed separately (tl_get_option
    "last_cg_phase_ms").  Iteration caps make the legs finish in seconds: both solvers stop on maxiters,
    the per-iteration figures do not depend on it."""
    import tealeaf_jl_b200 as tl
    from tealeaf_jl_b200 import dist as tld
    from tealeaf_jl_b200.device import DeviceChunk
    from tealeaf_jl_b200.solvers import get_solver
    t_begin = time.perf_counter()

    def solve_leg(solver, nx, ny, maxiters, over=None, options=None, tile_per_gpu=False):
        s = classic(nx, ny, 1, maxiters=maxiters, solver=solver)
        for k, v in (over or {}).items():
            setattr(s, k, v)
        if world > 1 and not tile_per_gpu:
            chunk, geom, _ = tld.create_tile(s, dist, local_rank, options=options)
        else:
            chunk, geom = tl.initialiseapp(s, backend=DeviceChunk, device=local_rank)
            for k, v in (options or {}).items():
                chunk.set_option(k, v)
        rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
        best = None
        for _ in range(2):                       # the first pass builds the graphs
            chunk.copy_field("energy", "energy0")
            tl.haloupdate(chunk, s, 1, ["energy", "density"])
            info = get_solver(solver).solve(chunk, s, rx, ry)
            info["cg_ms"] = chunk.get_option("last_cg_phase_ms")
            if best is None or info["solve_ms"] < best["solve_ms"]:
                best = info
        chunk.close()
        ms, cg_ms = maxr(best["solve_ms"]), maxr(best["cg_ms"])
        cells = nx * ny
        ngpu = 1 if tile_per_gpu else world
        out = {"global_cells": [nx, ny], "n_gpus": ngpu, "options": options or {}, "max_iters": maxiters,
               "iters": best["iters"], "cg_iters": best["cg_iters"], "solve_ms": ms, "cg_phase_ms": cg_ms,
               "error": best["error"], "kernel_launches": best["kernel_launches"]}
        if solver == "cg":
            sweeps, alg = best["iters"], 104 * best["iters"]
            out.update(us_per_iteration=1e3 * ms / max(best["iters"], 1))
        elif solver == "cheby":
            ch = best["cheby_iters"]
            sweeps, alg = best["iters"], 104 * best["cg_iters"] + 88 * ch
            ph = max(ms - cg_ms, 1e-9)
            out.update(cheby_iters=ch, cheby_phase_ms=ph, us_per_cheby_iteration=1e3 * ph / max(ch, 1),
                       cheby_cell_iterations_per_s=cells * ch / (ph * 1e-3),
                       cheby_algorithmic_gbs_per_gpu=88 * cells * ch / (ph * 1e-3) / 1e9 / ngpu,
                       cheby_frac_of_measured_peak=88 * cells * ch / (ph * 1e-3) / 1e9 / ngpu / peak)
        else:
            outer, inner = best["cheby_iters"], best["inner_total"]
            sweeps, alg = best["cg_iters"] + outer + inner, 104 * best["cg_iters"] + 128 * outer + 80 * inner
            ph = max(ms - cg_ms, 1e-9)
            out.update(outer_iters=outer, inner_steps_total=inner, halo_depth_k=best.get("halo_depth_k"), ppcg_phase_ms=ph,
                       us_per_inner_step_incl_outer=1e3 * ph / max(inner, 1), us_per_sweep=1e3 * ph / max(outer + inner, 1),
                       ppcg_cell_sweeps_per_s=cells * (outer + inner) / (ph * 1e-3),
                       ppcg_algorithmic_gbs_per_gpu=(128 * outer + 80 * inner) * cells / (ph * 1e-3) / 1e9 / ngpu,
                       ppcg_frac_of_measured_peak=(128 * outer + 80 * inner) * cells / (ph * 1e-3) / 1e9 / ngpu / peak)
        out.update(cell_iterations_per_s=cells * sweeps / (ms * 1e-3),
                   algorithmic_gbs_per_gpu=alg * cells / (ms * 1e-3) / 1e9 / ngpu,
                   frac_of_measured_peak=alg * cells / (ms * 1e-3) / 1e9 / ngpu / peak)
        return out

    legs = {"note": "BASELINE.json configs[2] (Chebyshev 4096^2 global) and configs[3] (PPCG 8192^2 global, 10 inner steps) solved on "
                    "the N GPUs of this run (strong scaling), iteration-capped, device-timed, max over ranks; algorithmic bytes per "
                    "cell: CG 104, Chebyshev 88, PPCG 128 per outer + 80 per inner step (SURVEY.md section 8d); fractions are of the "
                    "measured copy bandwidth per GPU.  default = two iterations / inner steps per pass (on tiles: one exchange per two)"}
    plan = [("chebyshev_4096", "cheby", 4096, 4096, 2600, None, None),
            ("ppcg_8192", "ppcg", 8192, 8192, 2600, None, None),
            ("chebyshev_4096_one_per_pass", "cheby", 4096, 4096, 2600, None, {"cheby_pair": 0}),
            ("ppcg_8192_one_per_pass_exchange_every_step", "ppcg", 8192, 8192, 2600, {"ppcghalodepth": 1}, {"ppcg_pair": 0})]
    if world > 1:
        plan.append(("ppcg_8192_one_per_pass_exchange_every_halo_depth", "ppcg", 8192, 8192, 2600, {"ppcghalodepth": 2}, {"ppcg_pair": 0}))
        plan.append(("cg_4096_strong", "cg", 4096, 4096, 1500, None, None))
    for name, solver, nx, ny, cap, over, opts in plan:
        # every rank must take the same decision: rank 0's clock decides
        go = 1.0 if time.perf_counter() - t_begin < budget_s else 0.0
        if dist is not None:
            import torch
            t = torch.tensor([go], dtype=torch.float64, device="cuda")
            dist.broadcast(t, 0)
            go = float(t.item())
        if not go:
            legs[name] = {"skipped": "time budget of the side legs spent"}
            continue
        try:
            legs[name] = solve_leg(solver, nx, ny, cap, over, opts)
        except Exception as e:       # a side leg must never break the headline line
            legs[name] = {"error": repr(e)[:300]}
            if dist is not None:
                break                # the ranks may no longer be in step
    if world == 1:
        try:
            legs["weak_tile_16384"] = dict(solve_leg("cg", 16384, 16384, 200, None, None, tile_per_gpu=True),
                                           note="the N > 1 workload's tile (16384^2 cells, CG capped at 200 iterations) solved alone: "
                                                "weak-scaling efficiency at N GPUs = value_N / (N x this cell_iterations_per_s)")
        except Exception as e:
            legs["weak_tile_16384"] = {"error": repr(e)[:300]}
    legs["seconds"] = time.perf_counter() - t_begin
    return legs


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
    solo = solo_fastest = None
    if world > 1:
        ssolo = classic(tile_nx, tile_ny, 1, maxiters=maxiters)     # the same iteration cap as the N-GPU timestep
        csolo = DeviceChunk(tile_nx, tile_ny, ssolo.halodepth, ssolo.maxiters, device=local_rank)
        rxs, rys = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2

        def solo_step():
            # one timestep exactly as the timed region below runs it: state upload aside, halo update + CG solve + solvefinished
            csolo.haloupdate(["energy", "density"], 1)
            info = csolo.cg_solve(ssolo, rxs, rys)
            csolo.solvefinished(True)
            return info

        for k in range(2):          # first pass warms up (graphs), second is timed on the device
            csolo.paint_states(s, geom)
            csolo.haloupdate(["density", "energy0", "energy"], 1)
            csolo.copy_field("energy", "energy0")
            csolo.timer_start()
            isolo = solo_step()
            solo_ms = csolo.timer_stop()
        # every rank times its own tile alone; like the N-GPU value, the line reports the SLOWEST GPU of the job
        t_solo = torch.tensor([solo_ms, solo_ms], dtype=torch.float64, device="cuda")
        t_fast = t_solo.clone()
        dist.all_reduce(t_solo, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_fast, op=dist.ReduceOp.MIN)
        solo = tile_nx * tile_ny * isolo["iters"] / (float(t_solo[0].item()) * 1e-3)
        solo_fastest = tile_nx * tile_ny * isolo["iters"] / (float(t_fast[0].item()) * 1e-3)
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
    # kernel A advances u every second launch (option cg_lazy_u, default): launches alternate between two costs, so the
    # average launch is the mean of the two (tl_time_kernel "cg_fused_w" = a launch with the u update, "_odd" = without)
    lazy_u = int(chunk.get_option("cg_u_mode")) == 2
    ka_upd_ms = chunk.time_kernel("cg_fused_w", 30)
    ka_odd_ms = chunk.time_kernel("cg_fused_w_odd", 30) if lazy_u else ka_upd_ms
    ka_ms = 0.5 * (ka_upd_ms + ka_odd_ms)
    kb_ms = chunk.time_kernel("cg_fused_r", 30)
    it_ms = solve_ms / max(iters, 1)
    a_phys = KERNEL_A_PHYS_BYTES_LAZY if lazy_u else KERNEL_A_PHYS_BYTES
    # ncu dram bytes per launch at this tile size (lazy: mean of the two kinds of launch), or None
    traffic = TRAFFIC.get(f"k_cg_fused_w{'_lazy' if lazy_u else ''}@{tile_nx}x{tile_ny}")
    roofline = {
        "bound": "hbm", "kernel": "k_cg_fused_w_ring<UM, S, MINB> (CG kernel A" + (", u advanced every second launch)" if lazy_u else ")"),
        "achieved": KERNEL_A_ALG_BYTES * tile_cells / (ka_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": KERNEL_A_ALG_BYTES * tile_cells / (ka_ms * 1e-3) / 1e9 / peak, "traffic": traffic,
        "peak_source": peak_src, "avg_launch_ms": ka_ms,
        "launch_ms_with_u_update": ka_upd_ms, "launch_ms_without_u_update": ka_odd_ms,
        "algorithmic_bytes_per_cell": KERNEL_A_ALG_BYTES, "physical_bytes_per_cell": a_phys,
        "physical_gbs": a_phys * tile_cells / (ka_ms * 1e-3) / 1e9,
        "other_kernels": {"k_cg_fused_r": {"avg_launch_ms": kb_ms,
                                           "achieved": KERNEL_B_ALG_BYTES * tile_cells / (kb_ms * 1e-3) / 1e9}},
        "iteration": {"ms": it_ms, "algorithmic_bytes_per_cell": CG_ALG_BYTES,
                      "achieved": CG_ALG_BYTES * tile_cells / (it_ms * 1e-3) / 1e9,
                      "frac_of_measured": CG_ALG_BYTES * tile_cells / (it_ms * 1e-3) / 1e9 / peak,
                      "frac_of_nominal_8TBs": CG_ALG_BYTES * tile_cells / (it_ms * 1e-3) / 1e9 / 8000.0},
    }

    # ---- the reduction-free iterations of configs[2] / configs[3] on the same tile (kernels timed alone) ----
    kernels_alone = None
    if world == 1:
        try:
            def per_it(name, its_per_launch):
                ms = min(chunk.time_kernel(name, 30) for _ in range(2))
                return 1e3 * ms / its_per_launch
            c1, c2 = per_it("cheby_fused", 1), per_it("cheby_pair", 2)
            p1, p2 = per_it("ppcg_inner", 1), per_it("ppcg_pair", 2)
            tj = TRAFFIC.get
            kernels_alone = {
                "note": "kernels of configs[2]/[3] timed alone on this 4096x4096 tile (CUDA events, 30 launches); two_per_pass = "
                        "temporal blocking (k_cheby_pair_ring / k_ppcg_pair_ring); physical_bytes_per_cell_pass = ncu dram bytes",
                "chebyshev_iteration": {"one_per_pass_us": c1, "two_per_pass_us": c2, "algorithmic_bytes_per_cell": 88,
                                        "cell_iterations_per_s": tile_cells / (c2 * 1e-6),
                                        "algorithmic_gbs": 88 * tile_cells / (c2 * 1e-6) / 1e9,
                                        "dram_bytes_per_launch_two_per_pass": tj("k_cheby_pair_ring@4096x4096"),
                                        "physical_gbs": (tj("k_cheby_pair_ring@4096x4096") or 0) / (2 * c2 * 1e-6) / 1e9},
                "ppcg_inner_step": {"one_per_pass_us": p1, "two_per_pass_us": p2, "algorithmic_bytes_per_cell": 80,
                                    "cell_steps_per_s": tile_cells / (p2 * 1e-6),
                                    "algorithmic_gbs": 80 * tile_cells / (p2 * 1e-6) / 1e9,
                                    "dram_bytes_per_launch_two_per_pass": tj("k_ppcg_pair_ring@4096x4096"),
                                    "physical_gbs": (tj("k_ppcg_pair_ring@4096x4096") or 0) / (2 * p2 * 1e-6) / 1e9},
            }
        except Exception as e:      # never let the side measurement break the headline line
            kernels_alone = {"error": repr(e)}
    chunk.close()
    chunk = None
    other = other_config_legs(world, rank, local_rank, dist, peak, maxr) if not args.no_legs else None
    if other is not None and kernels_alone is not None:
        other["kernels_alone_4096"] = kernels_alone

    cpu = cpu_serial = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.oracle import load
        threads = load().tlo_max_threads()
        v, secs = cpu_cg_sample(4096, 600, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"600 CG iterations of the same 4096x4096 deck ({secs:.1f} s), OpenMP oracle"}
        v1, secs1 = cpu_cg_sample(4096, 40, 1)      # what the reference is: serial (no threading anywhere, CG.jl:82-90)
        cpu_serial = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
                      "sample": f"40 CG iterations of the same deck on one thread ({secs1:.1f} s)"}
    elif rank == 0:
        cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "N>1: measured at N=1 only"}

    if rank == 0:
        _emit(json.dumps({
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
            **({"cpu_baseline_serial": cpu_serial} if cpu_serial else {}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * field_bytes,
                    "d2h_bytes_per_step": field_bytes + 32, "ms_per_step": 1e3 * e2e_wall / args.steps,
                    "host_numa_binding": numa},
            "gpu_launches": int(launches), "clocks": clocks,
            **({"other_configs": other} if other else {}),
            **({"same_tile_single_gpu": {"value": solo, "fastest_gpu_value": solo_fastest, "unit": UNIT, "note":
                "every rank's tile solved alone (1x1, one timestep exactly as timed at N GPUs: same iteration cap) in the same job, value = the slowest GPU's (as the N-GPU value is a max over ranks), fastest_gpu_value = the fastest one's; "
                "weak-scaling efficiency of the N-GPU workload = value / (N x this)"}} if solo else {}),
        }))
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
    ap.add_argument("--no-legs", action="store_true", help="skip other_configs (configs[2]/[3] and the weak base tile)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line.  Whatever else the process and its libraries print (NCCL's version banner
    # went to stdout in spite of NCCL_DEBUG_FILE on some boxes) is sent to stderr: file descriptor 1 is pointed at
    # stderr for the whole run and the JSON line is written to a private duplicate of the original stdout.
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if args.impl == "reference":
        from oracle import oracle as _o      # the CPU arm never builds or loads the CUDA library
        if int(os.environ.get("RANK", "0")) == 0:
            _o.build()
        run_reference(args)
    else:
        import __graft_entry__
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            __graft_entry__.build()
        run_b200(args)


if __name__ == "__main__":
    main()
