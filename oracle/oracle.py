"""Python wrapper of the CPU oracle (oracle/tealeaf_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product path.  PARITY UNPINNED against the
reference itself (no Julia here, no golden vectors in the reference), pinned externally by
upstream TeaLeaf's published QA checking values (tests/test_upstream_pin.py); see the C file header.

`OracleChunk` exposes the same kernel names as tealeaf.jl_b200/device.py's DeviceChunk, so the
host driver (app.diffuse, solvers.*.solve_stepwise) runs unchanged on either.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")

FIELD_NAMES = ["density", "energy0", "energy", "u", "u0", "p", "r", "w", "kx", "ky", "sd"]
FIELD_IDS = {n: i for i, n in enumerate(FIELD_NAMES)}


class Result(C.Structure):
    _fields_ = [("iters", C.c_int), ("cg_iters", C.c_int), ("cheby_iters", C.c_int), ("est_iters", C.c_int),
                ("inner_total", C.c_int), ("status", C.c_int), ("error", C.c_double), ("eigmin", C.c_double),
                ("eigmax", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "tealeaf_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "_build/liboracle.so"], check=True, capture_output=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB)
        D, I, P = C.c_double, C.c_int, C.c_void_p
        sig = {
            "tlo_create": (P, [I, I, I, I]), "tlo_destroy": (None, [P]),
            "tlo_field": (C.POINTER(D), [P, I]), "tlo_coef_array": (C.POINTER(D), [P, I]),
            "tlo_set_literal_smvp": (None, [P, I]), "tlo_set_threads": (None, [P, I]), "tlo_max_threads": (I, []),
            "tlo_get_scalar": (D, [P, I]), "tlo_set_eigs": (None, [P, D, D]),
            "tlo_smvp": (D, [P, I, I, I]),
            "tlo_halo_update": (None, [P, C.c_uint, I]),
            "tlo_cg_init": (D, [P, I, D, D, C.POINTER(I)]),
            "tlo_cg_calc_w": (D, [P]), "tlo_cg_calc_ur": (D, [P, D]), "tlo_cg_calc_p": (None, [P, D]),
            "tlo_copy_u": (None, [P]), "tlo_calc_residual": (None, [P]), "tlo_finalise": (None, [P]),
            "tlo_norm2": (D, [P, I]),
            "tlo_field_summary": (None, [P, D] + [C.POINTER(D)] * 4),
            "tlo_eigenvalues": (I, [P, I]), "tlo_cheby_coef": (None, [P, I]),
            "tlo_cheby_calc_iter": (I, [P, D, D]),
            "tlo_cheby_init_fields": (D, [P]), "tlo_cheby_iterate": (D, [P, I, I, D]),
            "tlo_ppcg_init_sd": (None, [P]), "tlo_ppcg_inner": (None, [P, I]),
            "tlo_cg_solve": (None, [P, I, D, D, D, I, C.POINTER(Result)]),
            "tlo_cheby_solve": (None, [P, I, D, D, D, I, I, D, I, C.POINTER(Result)]),
            "tlo_ppcg_solve": (None, [P, I, D, D, D, I, I, D, I, I, C.POINTER(Result)]),
            "tlo_jacobi_init": (None, [P, I, D, D, C.POINTER(I)]), "tlo_jacobi_iterate": (D, [P]),
            "tlo_jacobi_solve": (None, [P, I, D, D, D, I, C.POINTER(Result)]),
            "tlo_solve_finished": (None, [P, I]),
            "tlo_cg_fixed_iters": (D, [P, D, I]),
        }
        for n, (r, a) in sig.items():
            f = getattr(l, n)
            f.restype, f.argtypes = r, a
        _lib = l
    return _lib


class OracleChunk:
    """CPU restatement of the reference Chunk + kernels; same surface as DeviceChunk."""

    def __init__(self, xcells, ycells, halodepth=2, maxiters=10_000, threads=1, literal_smvp=False, **_):
        self._l = load()
        self.nx, self.ny, self.hd = xcells, ycells, halodepth
        self.x, self.y = xcells + 2 * halodepth, ycells + 2 * halodepth
        self.maxiters = maxiters
        self.c = self._l.tlo_create(xcells, ycells, halodepth, maxiters)
        self._l.tlo_set_threads(self.c, threads)
        self._l.tlo_set_literal_smvp(self.c, int(literal_smvp))
        self.cgalpha = np.ctypeslib.as_array(self._l.tlo_coef_array(self.c, 0), shape=(maxiters,))
        self.cgbeta = np.ctypeslib.as_array(self._l.tlo_coef_array(self.c, 1), shape=(maxiters,))
        self.chalpha = np.ctypeslib.as_array(self._l.tlo_coef_array(self.c, 2), shape=(maxiters,))
        self.chbeta = np.ctypeslib.as_array(self._l.tlo_coef_array(self.c, 3), shape=(maxiters,))

    def close(self):
        if getattr(self, "c", None):
            self._l.tlo_destroy(self.c)
            self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def view(self, name):
        """writable (x, y) Fortran-order view of the oracle's own storage"""
        p = self._l.tlo_field(self.c, FIELD_IDS[name])
        return np.ctypeslib.as_array(p, shape=(self.y, self.x)).T

    def set_field(self, name, arr):
        self.view(name)[...] = arr

    def get_field(self, name):
        return np.array(self.view(name), order="F", copy=True)

    def copy_field(self, dst, src):
        self.view(dst)[...] = self.view(src)

    def haloupdate(self, fields, depth=1):
        m = 0
        for f in fields:
            m |= 1 << FIELD_IDS[f]
        self._l.tlo_halo_update(self.c, m, depth)

    def cg_init(self, coef, rx, ry):
        st = C.c_int()
        v = self._l.tlo_cg_init(self.c, coef, rx, ry, C.byref(st))
        if st.value:
            raise ValueError(f"Coefficient {coef} is not valid")
        return v

    def cg_w(self):
        return self._l.tlo_cg_calc_w(self.c)

    def cg_ur(self, alpha):
        return self._l.tlo_cg_calc_ur(self.c, alpha)

    def cg_p(self, beta):
        self._l.tlo_cg_calc_p(self.c, beta)

    def copyu(self):
        self._l.tlo_copy_u(self.c)

    def residual(self):
        self._l.tlo_calc_residual(self.c)

    def finalise(self):
        self._l.tlo_finalise(self.c)

    def solvefinished(self, checkresult=True):
        self._l.tlo_solve_finished(self.c, int(checkresult))

    def norm2(self, field):
        return self._l.tlo_norm2(self.c, FIELD_IDS[field])

    def smvp(self, field, k, j):
        return self._l.tlo_smvp(self.c, FIELD_IDS[field], k, j)

    def cheby_init(self, theta):
        # theta lives in the oracle chunk; set it through eigs such that coef! reproduces it
        self._theta_override(theta)
        return self._l.tlo_cheby_init_fields(self.c)

    def _theta_override(self, theta):
        # tlo_cheby_coef sets theta = (eigmax+eigmin)/2; keep eigs if they already give theta
        if self._l.tlo_get_scalar(self.c, 0) != theta:
            self._l.tlo_set_eigs(self.c, theta * 0.5, theta * 1.5)
            self._l.tlo_cheby_coef(self.c, 0)
            assert self._l.tlo_get_scalar(self.c, 0) == theta

    def cheby_iterate(self, alpha, beta, calc2norm, error):
        self.chalpha[1], self.chbeta[1] = alpha, beta
        return self._l.tlo_cheby_iterate(self.c, 1, int(calc2norm), error)

    def ppcg_init_sd(self, theta):
        self._theta_override(theta)
        self._l.tlo_ppcg_init_sd(self.c)

    def ppcg_inner(self, alphas, betas, nsteps):
        self.chalpha[:nsteps] = alphas[:nsteps]
        self.chbeta[:nsteps] = betas[:nsteps]
        for pp in range(1, nsteps + 1):
            self._l.tlo_ppcg_inner(self.c, pp)

    def fieldsummary(self, cell_volume):
        v = [C.c_double() for _ in range(4)]
        self._l.tlo_field_summary(self.c, cell_volume, *[C.byref(q) for q in v])
        return tuple(q.value for q in v)

    def eigenvalues(self, cgiters):
        rc = self._l.tlo_eigenvalues(self.c, cgiters)
        return rc, self._l.tlo_get_scalar(self.c, 1), self._l.tlo_get_scalar(self.c, 2)

    def _res(self, r):
        d = r.as_dict()
        if d["status"]:
            raise RuntimeError(f"oracle solve failed: status {d['status']}")
        return d

    def cg_solve(self, s, rx, ry):
        r = Result()
        self._l.tlo_cg_solve(self.c, s.coefficient, rx, ry, s.eps, s.maxiters, C.byref(r))
        return self._res(r)

    def cheby_solve(self, s, rx, ry):
        r = Result()
        self._l.tlo_cheby_solve(self.c, s.coefficient, rx, ry, s.eps, s.maxiters, s.presteps, s.epslim,
                                int(s.errorswitch), C.byref(r))
        return self._res(r)

    def ppcg_solve(self, s, rx, ry):
        r = Result()
        self._l.tlo_ppcg_solve(self.c, s.coefficient, rx, ry, s.eps, s.maxiters, s.presteps, s.epslim,
                               int(s.errorswitch), s.ppcginnersteps, C.byref(r))
        return self._res(r)

    def jacobi_init(self, coef, rx, ry):
        st = C.c_int()
        self._l.tlo_jacobi_init(self.c, coef, rx, ry, C.byref(st))
        if st.value:
            raise ValueError(f"Coefficient {coef} is not valid")

    def jacobi_iterate(self):
        return self._l.tlo_jacobi_iterate(self.c)

    def jacobi_solve(self, s, rx, ry):
        r = Result()
        self._l.tlo_jacobi_solve(self.c, s.coefficient, rx, ry, s.eps, s.maxiters, C.byref(r))
        return self._res(r)

    def cg_fixed_iters(self, rro, iters):
        return self._l.tlo_cg_fixed_iters(self.c, rro, iters)
