"""Settings / deck parser -- host-side mirror of the reference's `src/settings.jl`.

Julia stays in charge of settings parsing in the real integration (julia/TeaLeafB200.jl);
this mirror exists because there is no Julia binary in the build/test environment, so the
parity tests and bench.py need the same `tea.in` / `tea.problems` grammar in Python.

Follows `src/settings.jl:39-78` (defaults), `:90-135` (`Settings(infile)`), `:142-173`
(`readstate`), `:180-196` (`checkingvalue`) including the quirks listed in SURVEY.md
Appendix B, with two documented corrections from Appendix A:
  #6  the state-geometry nudge uses the final dx, dy (two-pass);
  #26 blank / comment lines in tea.problems are tolerated.
And three extensions the north-star asks for: a leading `tl_` on any key is stripped,
`use_chebyshev` is accepted as `use_cheby`, and `ppcg_halo_depth=k` sets the depth of the tile
exchange of the PPCG inner steps (matrix-powers groups; results do not depend on it).
"""
from __future__ import annotations

import dataclasses
import logging
from dataclasses import dataclass, field
from typing import List

log = logging.getLogger("tealeaf")

CONDUCTIVITY = 1        # src/settings.jl:13
RECIP_CONDUCTIVITY = 2  # src/settings.jl:14

# src/settings.jl:11  @enum Geometry Rectangular Circular Point
RECTANGULAR, CIRCULAR, POINT = "rectangle", "circular", "point"

# src/chunk.jl:12
EXCHANGE_FIELDS = ["density", "p", "energy0", "energy", "u", "sd"]

SOLVERS = ("cg", "cheby", "ppcg", "jacobi")


@dataclass
class State:
    """src/settings.jl:20-29"""
    density: float = 0.0
    energy: float = 0.0
    xmin: float = 0.0
    ymin: float = 0.0
    xmax: float = 0.0
    ymax: float = 0.0
    radius: float = 0.0
    geometry: str = RECTANGULAR
    # raw (un-nudged) geometry as read; the nudge is applied once dx, dy are final (A#6)
    _raw: dict = field(default_factory=dict, repr=False)
    _num: int = 1


@dataclass
class Settings:
    """src/settings.jl:39-78 (same field names, same defaults)."""
    endstep: int = 2**63 - 1
    presteps: int = 30
    maxiters: int = 10_000
    coefficient: int = CONDUCTIVITY
    ppcginnersteps: int = 10
    summaryfrequency: int = 10
    halodepth: int = 2
    toexchange: dict = field(default_factory=lambda: {f: False for f in EXCHANGE_FIELDS})
    errorswitch: bool = False
    checkresult: bool = True
    eps: float = 1e-15
    dtinit: float = 0.1
    endtime: float = 10.0
    epslim: float = 1e-5
    solver: str = "cg"
    xcells: int = 10
    ycells: int = 10
    xmin: float = 0.0
    ymin: float = 0.0
    xmax: float = 100.0
    ymax: float = 100.0
    dx: float = 10.0
    dy: float = 10.0
    states: List[State] = field(default_factory=list)
    debugfile: str = ""
    problemfile: str = "tea.problems"
    # extension (north_star "depth-k halos"): exchange depth of the PPCG inner steps between tiles,
    # 0 = halo_depth.  Deck key `ppcg_halo_depth` / `tl_ppcg_halo_depth`; no effect on the result.
    ppcghalodepth: int = 0
    # compatibility switch (Appendix A #6): True reproduces the reference's literal state-geometry nudge,
    # which always uses the DEFAULT dx = dy = (100 - 0)/10 = 10 because `readstate` runs inside the line
    # loop and dx, dy are only recomputed after it (src/settings.jl:98-100 vs :132-133)
    literalnudge: bool = False

    def recompute_spacing(self) -> None:
        """src/settings.jl:132-133 (and Appendix A #23 for the -x/-y overrides)."""
        self.dx = (self.xmax - self.xmin) / self.xcells
        self.dy = (self.ymax - self.ymin) / self.ycells
        for s in self.states:
            _apply_nudge(s, self)


def resettoexchange(s: Settings) -> None:
    """src/settings.jl:80"""
    for f in EXCHANGE_FIELDS:
        s.toexchange[f] = False


_FIELD_TYPES = {f.name: f.type for f in dataclasses.fields(Settings)}
_SETTABLE = {
    "endstep": int, "presteps": int, "maxiters": int, "coefficient": int, "ppcginnersteps": int,
    "summaryfrequency": int, "halodepth": int, "xcells": int, "ycells": int, "ppcghalodepth": int,
    "errorswitch": bool, "checkresult": bool,
    "eps": float, "dtinit": float, "endtime": float, "epslim": float,
    "xmin": float, "ymin": float, "xmax": float, "ymax": float,
    "debugfile": str,
}


def _parse_value(ty, val: str):
    val = val.strip()
    if ty is bool:  # Julia parse(Bool, ..) accepts "true"/"false" (and 1/0)
        if val in ("true", "1"):
            return True
        if val in ("false", "0"):
            return False
        raise ValueError(val)
    if ty is int:
        return int(val)
    if ty is float:
        return float(val.replace("d", "e").replace("D", "e"))
    return val


def _apply_nudge(state: State, settings: Settings) -> None:
    """src/settings.jl:158-162: xmin/ymin + d/100, xmax/ymax - d/100 (state != 1 only)."""
    raw = state._raw
    if state._num == 1:
        return
    dx, dy = (10.0, 10.0) if settings.literalnudge else (settings.dx, settings.dy)
    if "xmin" in raw:
        state.xmin = raw["xmin"] + dx / 100
    if "ymin" in raw:
        state.ymin = raw["ymin"] + dy / 100
    if "xmax" in raw:
        state.xmax = raw["xmax"] - dx / 100
    if "ymax" in raw:
        state.ymax = raw["ymax"] - dy / 100


def readstate(line: str, settings: Settings) -> State:
    """src/settings.jl:142-173.  Single-space split, `key=val` tokens."""
    parts = line.split(" ")
    num = int(parts[1])
    state = State()
    state._num = num
    for pair in parts[2:]:
        if pair == "":
            # the reference would fail destructuring an empty token; be lenient with trailing blanks
            continue
        key, val = pair.split("=")
        if key == "density":
            state.density = float(val)
        elif key == "energy":
            state.energy = float(val)
        if num == 1:  # src/settings.jl:156: state 1 is the default state, geometry irrelevant
            continue
        if key in ("xmin", "ymin", "xmax", "ymax"):
            state._raw[key] = float(val)
        elif key == "radius":
            state.radius = float(val)
        elif key == "geometry":
            if val not in (RECTANGULAR, CIRCULAR, POINT):
                raise ValueError(f"unknown geometry {val!r}")
            state.geometry = val
    _apply_nudge(state, settings)
    return state


def parse_settings_text(text: str, literal_nudge: bool = False) -> Settings:
    """`Settings(infile)`, src/settings.jl:90-135, on the text of a deck.  `literal_nudge=True` paints
    the states exactly as the unpatched reference would (see Settings.literalnudge)."""
    settings = Settings()
    settings.literalnudge = literal_nudge
    state_lines = []
    for rawline in text.splitlines():
        line = rawline.strip()
        if line.startswith("state"):
            state_lines.append(line)
            continue
        if line.startswith("tl_"):  # extension: upstream-style tl_ prefix
            line = line[3:]
        if line.startswith("use_"):
            name = line[4:].strip()
            if name == "chebyshev":
                name = "cheby"
            if name in SOLVERS:
                settings.solver = name
            else:
                log.warning("Unknown setting %s", line)
            continue
        if line.startswith("*") or "=" not in line:
            continue
        key, val = line.split("=", 1)
        key = key.replace("_", "").strip()
        if key == "initialtimestep":
            key = "dtinit"
        ty = _SETTABLE.get(key)
        if ty is None:
            log.warning("Unknown setting %s (key %s)", line, key)
            continue
        try:
            setattr(settings, key, _parse_value(ty, val))
        except ValueError:
            log.warning("Unknown setting %s (key %s)", line, key)
    settings.dx = (settings.xmax - settings.xmin) / settings.xcells
    settings.dy = (settings.ymax - settings.ymin) / settings.ycells
    # second pass (Appendix A #6): states are read with the final dx, dy
    for line in state_lines:
        settings.states.append(readstate(line, settings))
    return settings


def parse_settings(infile: str, literal_nudge: bool = False) -> Settings:
    log.info("Reading configuration from %s", infile)
    with open(infile, "r") as fh:
        return parse_settings_text(fh.read(), literal_nudge=literal_nudge)


def checkingvalue(settings: Settings, problemfile: str | None = None) -> float:
    """src/settings.jl:180-196: `<x_cells> <y_cells> <end_step> <value>` per line."""
    problemfile = problemfile or settings.problemfile
    try:
        fh = open(problemfile, "r")
    except OSError:
        log.warning("Problem file %s not found.", problemfile)
        return 1.0
    with fh:
        for line in fh:
            parts = line.split()
            if len(parts) < 4 or parts[0].startswith(("*", "#")):
                continue
            try:
                params = [int(v) for v in parts[:3]]
                value = float(parts[3])
            except ValueError:
                continue
            if params == [settings.xcells, settings.ycells, settings.endstep]:
                return value
    log.warning("Problem was not found in the test problems file.")
    return 1.0
