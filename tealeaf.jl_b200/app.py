"""Application driver -- host-side mirror of `src/TeaLeaf.jl` (`initialiseapp!`, `diffuse!`).

In the real integration this loop stays in Julia (north-star: "Julia host code stays in
charge of settings parsing, the timestep loop and field summaries"); the mirror drives the
same sequence of calls through the same C-ABI so that parity tests and bench.py exercise
exactly what the Julia glue would.
"""
from __future__ import annotations

import logging

from .chunk import HostGeometry, paint_states
from .settings import Settings, checkingvalue
from .solvers import get_solver, haloupdate

log = logging.getLogger("tealeaf")


def upload_initial_state(chunk, settings: Settings, geom: HostGeometry | None = None, host_paint: bool = False):
    """The field part of `initialiseapp!` (src/TeaLeaf.jl:35-44) for an existing backend chunk.
    A backend that can paint the states itself (DeviceChunk: `tl_paint_states`, bit-identical to
    the host painter -- tests/test_gpu_parity.py) does so; otherwise (the oracle, or
    host_paint=True) the host paints and the three fields are uploaded."""
    geom = geom or HostGeometry(settings)
    if hasattr(chunk, "paint_states") and not host_paint:
        chunk.paint_states(settings, geom)                          # setchunkstate!, TeaLeaf.jl:37
    else:
        density, energy0, u = paint_states(settings, geom)
        chunk.set_field("density", density)
        chunk.set_field("energy0", energy0)
        chunk.set_field("u", u)
    haloupdate(chunk, settings, 1, ["density", "energy0", "energy"])  # TeaLeaf.jl:39
    chunk.copy_field("energy", "energy0")                       # TeaLeaf.jl:41
    return geom


def initialiseapp(settings: Settings, backend=None, **kw):
    """`initialiseapp!`, src/TeaLeaf.jl:35-44.  Returns (chunk, geometry)."""
    if backend is None:
        from .device import DeviceChunk
        backend = DeviceChunk
    chunk = backend(settings.xcells, settings.ycells, settings.halodepth, settings.maxiters, **kw)
    geom = upload_initial_state(chunk, settings)
    return chunk, geom


def fieldsummary(chunk, settings: Settings, geom: HostGeometry):
    """`fieldsummary`, src/kernels.jl:119-133 -> dict(vol, mass, ie, temp, qa_diff, passed)"""
    vol, mass, ie, temp = chunk.fieldsummary(geom.cell_volume)
    out = {"vol": vol, "mass": mass, "ie": ie, "temp": temp}
    if settings.checkresult:
        cv = checkingvalue(settings)
        qa_diff = abs(100.0 * temp / cv - 100.0)   # kernels.jl:127
        out.update(cv=cv, qa_diff=qa_diff, passed=qa_diff < 0.001)
        log.info("Checking results... cv=%r actual=%r", cv, temp)
        if qa_diff < 0.001:
            log.info("This run PASSED qa_diff=%r", qa_diff)
        else:
            log.warning("This run FAILED qa_diff=%r", qa_diff)
    return out


def julia_float(v: float) -> str:
    """`string(::Float64)` as Julia prints it (shortest round-trip digits; fixed notation for
    decimal exponents -5 < e < 6, else `d.ddde±x`), so that debug dumps diff against the reference's."""
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "Inf" if v > 0 else "-Inf"
    if v == 0.0:
        return "-0.0" if str(v).startswith("-") else "0.0"
    mant, _, exp = f"{v!r}".partition("e")
    sign = "-" if mant.startswith("-") else ""
    mant = mant.lstrip("-")
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0") or "0"
    e10 = (int(exp) if exp else 0) + len(ip.lstrip("0")) - 1 if ip.strip("0") else \
        (int(exp) if exp else 0) - (len(fp) - len(fp.lstrip("0"))) - 1
    digits = digits.rstrip("0") or "0"
    if -5 < e10 < 6:
        if e10 >= 0:
            whole, frac = digits[:e10 + 1].ljust(e10 + 1, "0"), digits[e10 + 1:]
            return f"{sign}{whole}.{frac or '0'}"
        return f"{sign}0.{'0' * (-e10 - 1)}{digits}"
    return f"{sign}{digits[0]}.{digits[1:] or '0'}e{e10}"


def debugrecord(settings: Settings, chunk, geom: HostGeometry):
    """`debugrecord`, src/TeaLeaf.jl:90-103: appends every `Chunk` attribute (src/chunk.jl:19-60, in
    declaration order, one matrix column per line) to `settings.debugfile`.  Device fields are
    downloaded with `get_field`; `density0`, `mi` (never written on the path) are zeros; `xarea`/
    `yarea` are skipped (Appendix A #5: never filled consistently, never read)."""
    if not settings.debugfile:
        return
    import numpy as np
    log.info("Writing debug data to %s", settings.debugfile)
    zeros = np.zeros((geom.x, geom.y))

    def mat(a):
        return "\n".join(" ".join(julia_float(float(v)) for v in a[:, j]) for j in range(a.shape[1]))

    def vec(a):
        return " ".join(julia_float(float(v)) for v in a)

    n = getattr(chunk, "maxiters", 0)
    pick = lambda name: vec(np.asarray(getattr(chunk, name, np.zeros(n)))[:n])
    items = [("density0", mat(zeros)), ("density", mat(chunk.get_field("density"))),
             ("energy0", mat(chunk.get_field("energy0"))), ("energy", mat(chunk.get_field("energy")))]
    items += [(f, mat(chunk.get_field(f))) for f in ("u", "u0", "p", "r")]
    items += [("mi", mat(zeros))] + [(f, mat(chunk.get_field(f))) for f in ("w", "kx", "ky", "sd")]
    items += [("vertexx", vec(geom.vertexx)), ("vertexy", vec(geom.vertexy)), ("cellx", vec(geom.cellx)),
              ("celly", vec(geom.celly)), ("volume", mat(zeros + geom.cell_volume))]
    items += [("θ", julia_float(getattr(chunk, "theta", 0.0))), ("eigmin", julia_float(getattr(chunk, "eigmin", 0.0))),
              ("eigmax", julia_float(getattr(chunk, "eigmax", 0.0)))]
    items += [("cgα", pick("cgalpha")), ("cgβ", pick("cgbeta")), ("chebyα", pick("chalpha")), ("chebyβ", pick("chbeta"))]
    with open(settings.debugfile, "a", encoding="utf-8") as fh:
        for name, text in items:
            fh.write(f"{name}\n{text}\n\n")
        fh.write("\n\n")


def write_tea_out(path, settings: Settings, records, final, wall_s=None):
    """`tea.out`-style report: the reference only logs (`@info`, src/TeaLeaf.jl:80,
    src/kernels.jl:125-131); upstream TeaLeaf writes this table.  One block per timestep (solver, iterations,
    error), the four-component field summary where one was taken, the QA verdict at the end."""
    head = f"{'':>12}{'Volume':>16}{'Mass':>16}{'Density':>16}{'Energy':>16}{'U':>16}"

    def row(tag, s):
        dens = s["mass"] / s["vol"] if s["vol"] else 0.0
        return f"{tag:>12}{s['vol']:16.7E}{s['mass']:16.7E}{dens:16.7E}{s['ie']:16.7E}{s['temp']:16.7E}"

    lines = ["Tea Version libtealeaf_b200", f" Mesh {settings.xcells} x {settings.ycells}, solver {settings.solver}, "
             f"dt {settings.dtinit!r}, eps {settings.eps!r}, max_iters {settings.maxiters}", ""]
    for r in records:
        lines.append(f" Step {r['step']:7d} time {r['step'] * settings.dtinit:.7E} timestep {settings.dtinit:.7E}")
        extra = "".join(f", {k} {r[k]}" for k in ("cheby_iters", "inner_total", "est_iters") if r.get(k))
        lines.append(f" Conduction error {r.get('error', 0.0):.7E}")
        lines.append(f" Iteration count {r.get('iters', 0):8d}{extra}")
        if "summary" in r:
            lines += ["", head, row(f"step:{r['step']:7d}", r["summary"]), ""]
    lines += ["", head, row("final:", final), ""]
    if "cv" in final:
        lines.append(f" Checking results... expected {final['cv']:.15E} got {final['temp']:.15E} "
                     f"diff {final['qa_diff']:.7E} %")
        lines.append(" This test is considered PASSED" if final["passed"] else " This test is considered NOT PASSED")
    if wall_s is not None:
        lines.append(f" Wall clock {wall_s:.6f} s")
    with open(path, "w", encoding="utf-8") as fh:
        fh.write("\n".join(lines) + "\n")


def diffuse(chunk, settings: Settings, geom: HostGeometry, stepwise: bool = False, on_step=None):
    """`diffuse!`, src/TeaLeaf.jl:62-83.  Returns the per-step records."""
    if settings.endstep >= 2**62:
        raise ValueError("end_step is required (SURVEY Appendix A #22)")
    solver = get_solver(settings.solver)
    records = []
    import os
    if settings.debugfile and os.path.isfile(settings.debugfile):   # TeaLeaf.jl:63-65
        os.remove(settings.debugfile)
    for tt in range(1, settings.endstep + 1):
        debugrecord(settings, chunk, geom)                          # TeaLeaf.jl:68
        rx = settings.dtinit / settings.dx ** 2    # TeaLeaf.jl:69
        ry = settings.dtinit / settings.dy ** 2    # TeaLeaf.jl:70
        haloupdate(chunk, settings, 1, ["energy", "density"])   # TeaLeaf.jl:71
        info = (solver.solve_stepwise if stepwise else solver.solve)(chunk, settings, rx, ry)  # TeaLeaf.jl:74
        chunk.solvefinished(settings.checkresult)                # TeaLeaf.jl:77
        # solvefinished! ends with haloupdate!(.., [:energy], false): marks energy, no reset
        settings.toexchange["energy"] = True
        rec = {"step": tt, **info}
        if tt % settings.summaryfrequency == 0:                  # TeaLeaf.jl:79
            rec["summary"] = fieldsummary(chunk, settings, geom)
        log.info("Timestep %d finished", tt)
        records.append(rec)
        if on_step:
            on_step(rec)
        # end_time: parsed by the reference (settings.jl:58) and never used; upstream's rule -- the loop also ends
        # once the simulated time reaches it.  The timestep is constant (TeaLeaf.jl:69-70).
        if tt * settings.dtinit >= settings.endtime:
            break
    final = fieldsummary(chunk, settings, geom)                  # TeaLeaf.jl:82
    return records, final
