"""Solver modules -- host-side mirror of `src/solvers/{CG,Cheby,PPCG,Jacobi}.jl`.

Two ways to run a solve, both behind the reference's entry point
`settings.solver.solve!(chunk, settings, rx, ry)` (src/TeaLeaf.jl:74):

* `solve(...)`           one call into the fused, graph-launched device path
                         (`tl_cg_solve` / `tl_cheby_solve` / `tl_ppcg_solve`);
* `solve_stepwise(...)`  the reference's own control flow, kernel by kernel, through the
                         per-function ABI (`tl_cg_calc_w`, ...).  This is what a Julia host
                         that keeps `solve!` written in Julia would execute, and it is how the
                         tests pin every individual kernel against the oracle.

Both accept any backend exposing the kernel names (DeviceChunk, or the oracle wrapper in
tests).  Corrections from SURVEY.md Appendix A are marked A#n.
"""
from __future__ import annotations

import math

import numpy as np

from .settings import Settings, resettoexchange

ERROR_START = 1e10        # src/kernels.jl:7
ERROR_SWITCH_MAX = 1.0    # src/kernels.jl:8
CGEIGENITERS = 20         # src/solvers/Cheby.jl:7


def haloupdate(chunk, settings: Settings, depth: int, toex=None, reset: bool = True):
    """`haloupdate!`, src/kernels.jl:146-159, with the sticky `settings.toexchange` set."""
    if toex is not None:
        if reset:
            resettoexchange(settings)
        for f in toex:
            settings.toexchange[f] = True
        fields = list(toex)
    else:
        fields = [f for f, on in settings.toexchange.items() if on]
    if fields:
        chunk.haloupdate(fields, depth)


# ------------------------------------------------------------------------------------------
# host-side scalar work (tiny; stays on the host in the reference as well)
# ------------------------------------------------------------------------------------------
def eigenvalues(cgalpha, cgbeta, cgiters: int):
    """`eigenvalues!`, src/kernels.jl:19-51 (A#9-11: a correct symmetric-tridiagonal solve)."""
    if cgiters < 1:
        raise RuntimeError("eigenvalues!: no CG iterations")
    a = np.asarray(cgalpha[:cgiters], dtype=np.float64)
    b = np.asarray(cgbeta[:cgiters], dtype=np.float64)
    diag = 1.0 / a
    diag[1:] += b[:-1] / a[:-1]
    off = np.sqrt(b[:-1]) / a[:-1]
    T = np.diag(diag) + np.diag(off, 1) + np.diag(off, -1)
    ev = np.linalg.eigvalsh(T)
    eigmin, eigmax = float(ev[0]), float(ev[-1])
    if eigmin < 0 or eigmax < 0:
        raise RuntimeError(f"Negative eigenvalue found: ({eigmin}, {eigmax})")  # kernels.jl:41-43
    return eigmin * 0.95, eigmax * 1.05


def cheby_coef(eigmin: float, eigmax: float, n: int):
    """`Cheby.coef!`, src/solvers/Cheby.jl:121-135 -> (theta, alphas, betas)"""
    theta = (eigmax + eigmin) / 2
    delta = (eigmax - eigmin) / 2
    sigma = theta / delta
    rho_old = 1 / sigma
    al, be = np.zeros(n), np.zeros(n)
    for i in range(n):
        rho_new = 1 / (2 * sigma - rho_old)
        al[i] = rho_new * rho_old
        be[i] = 2 * rho_new / delta
        rho_old = rho_new
    return theta, al, be


def cheby_calciter(eigmin: float, eigmax: float, error: float, bb: float) -> int:
    """`Cheby.calciter`, src/solvers/Cheby.jl:109-118"""
    connum = eigmax / eigmin
    it_alpha = np.finfo(np.float64).eps * bb / (4 * error)
    gamma = (math.sqrt(connum) - 1) / (math.sqrt(connum) + 1)
    return int(np.rint(math.log(it_alpha) / (2 * math.log(gamma))))


def _switch(settings: Settings, error: float, tt: int) -> bool:
    """Cheby.jl:25-29 / PPCG.jl:25-30"""
    if settings.errorswitch:
        return error < settings.epslim and tt > CGEIGENITERS
    return tt > settings.presteps and error < ERROR_SWITCH_MAX


# ------------------------------------------------------------------------------------------
# CG
# ------------------------------------------------------------------------------------------
class CG:
    name = "cg"

    @staticmethod
    def mainstep(chunk, settings: Settings, tt: int, rro: float) -> float:
        """src/solvers/CG.jl:32-44"""
        pw = chunk.cg_w()
        alpha = rro / pw
        chunk.cgalpha[tt - 1] = alpha
        rrn = chunk.cg_ur(alpha)
        beta = rrn / rro
        chunk.cgbeta[tt - 1] = beta
        chunk.cg_p(beta)
        return rrn

    @staticmethod
    def solve_stepwise(chunk, settings: Settings, rx: float, ry: float) -> dict:
        """src/solvers/CG.jl:7-29"""
        rro = chunk.cg_init(settings.coefficient, rx, ry)
        haloupdate(chunk, settings, 1, ["u", "p"])
        chunk.copyu()
        error = ERROR_START
        iters = 0
        for tt in range(1, settings.maxiters + 1):
            iters = tt
            rro = error = CG.mainstep(chunk, settings, tt, rro)
            haloupdate(chunk, settings, 1)
            if math.sqrt(abs(error)) < settings.eps:
                break
        return {"iters": iters, "cg_iters": iters, "error": error}

    @staticmethod
    def solve(chunk, settings: Settings, rx: float, ry: float) -> dict:
        info = chunk.cg_solve(settings, rx, ry)
        resettoexchange(settings)
        settings.toexchange["u"] = settings.toexchange["p"] = True
        return info


# ------------------------------------------------------------------------------------------
# Chebyshev
# ------------------------------------------------------------------------------------------
class Cheby:
    name = "cheby"

    @staticmethod
    def solve_stepwise(chunk, settings: Settings, rx: float, ry: float) -> dict:
        """src/solvers/Cheby.jl:10-61 with A#12, #13, #14"""
        error = ERROR_START
        estiter = 0
        chebyiters = 0
        cgiters = 0
        eigmin = eigmax = 0.0
        al = be = None
        rro = chunk.cg_init(settings.coefficient, rx, ry)
        haloupdate(chunk, settings, 1, ["u", "p"])  # A#12
        chunk.copyu()
        tt = 0
        for tt in range(1, settings.maxiters + 1):
            switch = chebyiters != 0 or _switch(settings, error, tt)
            if not switch:
                rro = error = CG.mainstep(chunk, settings, tt, rro)  # A#13
                cgiters = tt
            else:
                chebyiters += 1
                if chebyiters == 1:
                    eigmin, eigmax = eigenvalues(chunk.cgalpha, chunk.cgbeta, tt - 1)  # A#14
                    theta, al, be = cheby_coef(eigmin, eigmax, max(settings.maxiters - (tt - 1), 2))
                    # the reference indexes chebyα[chebyiters+1] in a maxiters-long, zero-initialised vector
                    # (chunk.jl:86-87): entries past the filled ones read 0.0, as tl_cheby_solve's padding does
                    al, be = np.concatenate([al, np.zeros(2)]), np.concatenate([be, np.zeros(2)])
                    bb = chunk.cheby_init(theta)
                    resettoexchange(settings)
                    settings.toexchange["u"] = True
                    error = chunk.cheby_iterate(al[1], be[1], True, error)
                    estiter = cheby_calciter(eigmin, eigmax, error, bb)
                else:
                    calc2norm = chebyiters >= estiter and (tt + 1) % 10 == 0
                    error = chunk.cheby_iterate(al[chebyiters], be[chebyiters], calc2norm, error)
            haloupdate(chunk, settings, 1)
            if abs(error) < settings.eps:
                break
        return {"iters": tt, "cg_iters": cgiters, "cheby_iters": chebyiters, "est_iters": estiter,
                "error": error, "eigmin": eigmin, "eigmax": eigmax}

    @staticmethod
    def solve(chunk, settings: Settings, rx: float, ry: float) -> dict:
        info = chunk.cheby_solve(settings, rx, ry)
        resettoexchange(settings)
        settings.toexchange["u"] = True
        return info


# ------------------------------------------------------------------------------------------
# PPCG
# ------------------------------------------------------------------------------------------
class PPCG:
    name = "ppcg"

    @staticmethod
    def mainstep(chunk, settings: Settings, rro: float, theta: float, al, be) -> float:
        """src/solvers/PPCG.jl:64-94 with A#16, #17, #18, #20"""
        pw = chunk.cg_w()
        alpha = rro / pw
        chunk.cg_ur(alpha)
        chunk.ppcg_init_sd(theta)                       # A#18
        resettoexchange(settings)
        settings.toexchange["sd"] = True
        chunk.ppcg_inner(al, be, settings.ppcginnersteps)  # halo(sd) + two-phase step, A#17
        resettoexchange(settings)
        settings.toexchange["p"] = True
        rrn = chunk.norm2("r")                          # A#20
        beta = rrn / rro
        chunk.cg_p(beta)
        return rrn

    @staticmethod
    def solve_stepwise(chunk, settings: Settings, rx: float, ry: float) -> dict:
        """src/solvers/PPCG.jl:9-55 with A#12, #14, #15, #16, #19"""
        ppcgiters = 0
        cgiters = 0
        eigmin = eigmax = theta = 0.0
        al = be = None
        rro = chunk.cg_init(settings.coefficient, rx, ry)
        haloupdate(chunk, settings, 1, ["u", "p"])  # A#12
        chunk.copyu()
        error = ERROR_START
        tt = 0
        for tt in range(1, settings.maxiters + 1):
            switch = ppcgiters != 0 or _switch(settings, error, tt)
            if not switch:
                rro = error = CG.mainstep(chunk, settings, tt, rro)
                cgiters = tt
            else:
                ppcgiters += 1  # A#15
                if ppcgiters == 1:
                    eigmin, eigmax = eigenvalues(chunk.cgalpha, chunk.cgbeta, tt - 1)  # A#14
                    theta, al, be = cheby_coef(eigmin, eigmax, settings.ppcginnersteps)
                    chunk.residual()                               # PPCG.jl:59
                    haloupdate(chunk, settings, 1, ["p"])          # PPCG.jl:60
                    rro = chunk.norm2("r")                         # A#19
                rro = error = PPCG.mainstep(chunk, settings, rro, theta, al, be)
            haloupdate(chunk, settings, 1)
            if abs(error) < settings.eps:
                break
        return {"iters": tt, "cg_iters": cgiters, "cheby_iters": ppcgiters,
                "inner_total": ppcgiters * settings.ppcginnersteps, "error": error,
                "eigmin": eigmin, "eigmax": eigmax}

    @staticmethod
    def solve(chunk, settings: Settings, rx: float, ry: float) -> dict:
        info = chunk.ppcg_solve(settings, rx, ry)
        resettoexchange(settings)
        settings.toexchange["p"] = True
        return info


# ------------------------------------------------------------------------------------------
# Jacobi  
# ------------------------------------------------------------------------------------------
class Jacobi:
    name = "jacobi"

    @staticmethod
    def solve_stepwise(chunk, settings: Settings, rx: float, ry: float) -> dict:
        """`Jacobi.driver!`, src/solvers/Jacobi.jl:7-31 (A#21: it is the module's `solve!`, and
        `error +=` is read as `error =`)."""
        chunk.jacobi_init(settings.coefficient, rx, ry)     # Jacobi.jl:8
        resettoexchange(settings)
        settings.toexchange["u"] = True                     # Jacobi.jl:55-56
        error = ERROR_START
        iters = 0
        for tt in range(1, settings.maxiters + 1):
            iters = tt
            error = chunk.jacobi_iterate()                  # Jacobi.jl:14
            if tt % 50 == 0:                                # Jacobi.jl:16-21
                haloupdate(chunk, settings, 1)
                chunk.residual()
                error = chunk.norm2("r")
            haloupdate(chunk, settings, 1)                  # Jacobi.jl:23
            if abs(error) < settings.eps:                   # Jacobi.jl:26
                break
        return {"iters": iters, "cg_iters": 0, "error": error}

    @staticmethod
    def solve(chunk, settings: Settings, rx: float, ry: float) -> dict:
        info = chunk.jacobi_solve(settings, rx, ry)
        resettoexchange(settings)
        settings.toexchange["u"] = True
        return info


SOLVER_MODULES = {"cg": CG, "cheby": Cheby, "ppcg": PPCG, "jacobi": Jacobi}


def get_solver(name: str):
    try:
        return SOLVER_MODULES[name]
    except KeyError:
        raise NotImplementedError(f"solver {name!r} is outside the hot-path scope (SURVEY.md §8f)") from None
