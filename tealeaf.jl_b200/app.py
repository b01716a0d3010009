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


def diffuse(chunk, settings: Settings, geom: HostGeometry, stepwise: bool = False, on_step=None):
    """`diffuse!`, src/TeaLeaf.jl:62-83.  Returns the per-step records."""
    if settings.endstep >= 2**62:
        raise ValueError("end_step is required (SURVEY Appendix A #22)")
    solver = get_solver(settings.solver)
    records = []
    for tt in range(1, settings.endstep + 1):
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
    final = fieldsummary(chunk, settings, geom)                  # TeaLeaf.jl:82
    return records, final
