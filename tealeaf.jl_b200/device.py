"""DeviceChunk -- the device-resident `Chunk` (src/chunk.jl:19-60) behind the C-ABI.

Method names follow the reference's kernel names (src/kernels.jl, src/solvers/*.jl); each
is a thin call into libtealeaf_b200.so.  Field data lives in HBM and only crosses the
boundary through set_field / get_field.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib
from .chunk import FIELD_IDS
from .settings import Settings


def _mask(fields) -> int:
    m = 0
    for f in fields:
        m |= 1 << FIELD_IDS[f]
    return m


class DeviceChunk:
    def __init__(self, xcells: int, ycells: int, halodepth: int = 2, maxiters: int = 10_000, device: int = 0,
                 rank: int = 0, px: int = 1, py: int = 1):
        self._l = _lib.load()
        self.nx, self.ny, self.hd = xcells, ycells, halodepth
        self.x, self.y = xcells + 2 * halodepth, ycells + 2 * halodepth
        self.maxiters = maxiters
        self.rank, self.px, self.py = rank, px, py
        ctx = C.c_void_p()
        rc = self._l.tl_create_tile(C.byref(ctx), xcells, ycells, halodepth, maxiters, device, rank, px, py)
        if rc != _lib.TL_OK:
            raise _lib.TeaLeafError(rc, "tl_create_tile failed (a B200 / sm_100 GPU is required; no CPU fallback)")
        self.ctx = ctx
        self.cgalpha = np.zeros(maxiters)  # chunk.cgα / cgβ, src/chunk.jl:56-57
        self.cgbeta = np.zeros(maxiters)

    @classmethod
    def multi(cls, xcells: int, ycells: int, halodepth: int = 2, maxiters: int = 10_000, ngpus: int = 2, devices=None,
              px: int = 0, py: int = 0, **_):
        """ONE chunk of the global mesh spread over `ngpus` GPUs of this process (`tl_create_multi`): same methods,
        `set_field` / `get_field` scatter / gather the global arrays.  `devices` may repeat an index (tiles sharing
        one GPU: test mode, needs CUDA_MODULE_LOADING=EAGER)."""
        self = cls.__new__(cls)
        self._l = _lib.load()
        self.nx, self.ny, self.hd = xcells, ycells, halodepth
        self.x, self.y = xcells + 2 * halodepth, ycells + 2 * halodepth
        self.maxiters = maxiters
        self.rank, self.px, self.py = 0, px, py
        ctx = C.c_void_p()
        dev = (C.c_int * ngpus)(*devices) if devices is not None else None
        rc = self._l.tl_create_multi(C.byref(ctx), xcells, ycells, halodepth, maxiters, ngpus, dev, px, py)
        if rc != _lib.TL_OK:
            raw = self._l.tl_last_error(None)
            raise _lib.TeaLeafError(rc, "tl_create_multi failed: " + (raw.decode() if raw else ""))
        self.ctx = ctx
        self.cgalpha = np.zeros(maxiters)
        self.cgbeta = np.zeros(maxiters)
        return self

    # ---- lifetime ----
    def close(self):
        if getattr(self, "ctx", None):
            self._l.tl_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _lib.check(self.ctx, rc)

    def set_option(self, name: str, value: float):
        self._ck(self._l.tl_set_option(self.ctx, name.encode(), float(value)))

    def get_option(self, name: str) -> float:
        out = C.c_double()
        self._ck(self._l.tl_get_option(self.ctx, name.encode(), C.byref(out)))
        return out.value

    # ---- multi-GPU wiring ----
    def comm_export(self) -> bytes:
        buf = C.create_string_buffer(self._l.tl_comm_blob_size())
        self._ck(self._l.tl_comm_export(self.ctx, buf))
        return buf.raw

    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._ck(self._l.tl_comm_unique_id(buf))
        return buf.raw

    def comm_connect(self, blobs, nccl_id: bytes | None = None):
        """Wire this tile to the others.  `nccl_id` (from rank 0's `comm_unique_id`) is only needed
        for the legacy mode (option comm_fused = 0); None gives a fused-only context."""
        allb = b"".join(blobs)
        self._ck(self._l.tl_comm_connect(self.ctx, C.c_char_p(allb), C.c_char_p(nccl_id) if nccl_id else None))

    # ---- field transfer ----
    def set_field(self, name: str, arr: np.ndarray):
        a = np.asfortranarray(arr, dtype=np.float64)
        if a.shape != (self.x, self.y):
            raise ValueError(f"field {name}: shape {a.shape} != {(self.x, self.y)}")
        self._ck(self._l.tl_set_field(self.ctx, FIELD_IDS[name], a.ctypes.data_as(C.c_void_p), self.x))

    def get_field(self, name: str) -> np.ndarray:
        a = np.empty((self.x, self.y), dtype=np.float64, order="F")
        self._ck(self._l.tl_get_field(self.ctx, FIELD_IDS[name], a.ctypes.data_as(C.c_void_p), self.x))
        return a

    def set_field_raw(self, name: str, ptr: int, ld: int):
        """pointer variant (pinned torch buffers in bench.py)"""
        self._ck(self._l.tl_set_field(self.ctx, FIELD_IDS[name], C.c_void_p(ptr), ld))

    def get_field_raw(self, name: str, ptr: int, ld: int):
        self._ck(self._l.tl_get_field(self.ctx, FIELD_IDS[name], C.c_void_p(ptr), ld))

    def paint_states(self, settings: Settings, geom) -> None:
        """`setchunkstate!` (src/chunk.jl:122-151) on the device: density, energy0, u."""
        from .settings import CIRCULAR, POINT, RECTANGULAR
        gid = {RECTANGULAR: 0, CIRCULAR: 1, POINT: 2}
        arr = (_lib.PaintState * len(settings.states))()
        for q, st in enumerate(settings.states):
            arr[q] = _lib.PaintState(st.density, st.energy, st.xmin, st.ymin, st.xmax, st.ymax, st.radius,
                                     gid[st.geometry], 0)
        self._ck(self._l.tl_paint_states(self.ctx, len(settings.states), arr, settings.xmin, settings.ymin,
                                         settings.dx, settings.dy, geom.x0, geom.y0))

    def copy_field(self, dst: str, src: str):
        self._ck(self._l.tl_copy_field(self.ctx, FIELD_IDS[dst], FIELD_IDS[src]))

    # ---- kernels (reference names) ----
    def haloupdate(self, fields, depth: int = 1):
        self._ck(self._l.tl_halo_update(self.ctx, _mask(fields), depth))

    def cg_init(self, coef: int, rx: float, ry: float) -> float:
        out = C.c_double()
        self._ck(self._l.tl_cg_init(self.ctx, coef, rx, ry, C.byref(out)))
        return out.value

    def cg_w(self) -> float:
        out = C.c_double()
        self._ck(self._l.tl_cg_calc_w(self.ctx, C.byref(out)))
        return out.value

    def cg_ur(self, alpha: float) -> float:
        out = C.c_double()
        self._ck(self._l.tl_cg_calc_ur(self.ctx, alpha, C.byref(out)))
        return out.value

    def cg_p(self, beta: float):
        self._ck(self._l.tl_cg_calc_p(self.ctx, beta))

    def copyu(self):
        self._ck(self._l.tl_copy_u(self.ctx))

    def residual(self):
        self._ck(self._l.tl_calc_residual(self.ctx))

    def finalise(self):
        self._ck(self._l.tl_finalise(self.ctx))

    def solvefinished(self, checkresult: bool = True):
        self._ck(self._l.tl_solve_finished(self.ctx, int(checkresult)))

    def norm2(self, field: str) -> float:
        out = C.c_double()
        self._ck(self._l.tl_norm2(self.ctx, FIELD_IDS[field], C.byref(out)))
        return out.value

    def cheby_init(self, theta: float) -> float:
        out = C.c_double()
        self._ck(self._l.tl_cheby_init(self.ctx, theta, C.byref(out)))
        return out.value

    def cheby_iterate(self, alpha: float, beta: float, calc2norm: bool, error: float) -> float:
        out = C.c_double(error)
        self._ck(self._l.tl_cheby_iterate(self.ctx, alpha, beta, int(calc2norm), C.byref(out)))
        return out.value

    def ppcg_init_sd(self, theta: float):
        self._ck(self._l.tl_ppcg_init_sd(self.ctx, theta))

    def ppcg_inner(self, alphas, betas, nsteps: int):
        a = np.ascontiguousarray(alphas, dtype=np.float64)
        b = np.ascontiguousarray(betas, dtype=np.float64)
        self._ck(self._l.tl_ppcg_inner(self.ctx, a.ctypes.data_as(C.POINTER(C.c_double)),
                                       b.ctypes.data_as(C.POINTER(C.c_double)), nsteps))

    def jacobi_init(self, coef: int, rx: float, ry: float):
        self._ck(self._l.tl_jacobi_init(self.ctx, coef, rx, ry))

    def jacobi_iterate(self) -> float:
        out = C.c_double()
        self._ck(self._l.tl_jacobi_iterate(self.ctx, C.byref(out)))
        return out.value

    def fieldsummary(self, cell_volume: float):
        v = [C.c_double() for _ in range(4)]
        self._ck(self._l.tl_field_summary(self.ctx, cell_volume, *[C.byref(q) for q in v]))
        return tuple(q.value for q in v)  # vol, mass, ie, temp

    # ---- whole-solve fast paths ----
    def cg_solve(self, s: Settings, rx: float, ry: float) -> dict:
        info = _lib.SolveInfo()
        self._ck(self._l.tl_cg_solve(self.ctx, s.coefficient, rx, ry, s.eps, min(s.maxiters, self.maxiters),
                                     C.byref(info), self.cgalpha.ctypes.data_as(C.POINTER(C.c_double)),
                                     self.cgbeta.ctypes.data_as(C.POINTER(C.c_double))))
        return info.as_dict()

    def cheby_solve(self, s: Settings, rx: float, ry: float) -> dict:
        info = _lib.SolveInfo()
        self._ck(self._l.tl_cheby_solve(self.ctx, s.coefficient, rx, ry, s.eps, min(s.maxiters, self.maxiters),
                                        s.presteps, s.epslim, int(s.errorswitch), C.byref(info)))
        return info.as_dict()

    def ppcg_solve(self, s: Settings, rx: float, ry: float) -> dict:
        info = _lib.SolveInfo()
        self._ck(self._l.tl_ppcg_solve(self.ctx, s.coefficient, rx, ry, s.eps, min(s.maxiters, self.maxiters),
                                       s.presteps, s.epslim, int(s.errorswitch), s.ppcginnersteps,
                                       int(s.ppcghalodepth), C.byref(info)))
        return info.as_dict()

    def jacobi_solve(self, s: Settings, rx: float, ry: float) -> dict:
        info = _lib.SolveInfo()
        self._ck(self._l.tl_jacobi_solve(self.ctx, s.coefficient, rx, ry, s.eps, min(s.maxiters, self.maxiters),
                                         C.byref(info)))
        return info.as_dict()

    def time_kernel(self, kernel: str, reps: int = 20) -> float:
        out = C.c_double()
        self._ck(self._l.tl_time_kernel(self.ctx, kernel.encode(), reps, C.byref(out)))
        return out.value

    def timer_start(self):
        self._ck(self._l.tl_timer_start(self.ctx))

    def timer_stop(self) -> float:
        out = C.c_double()
        self._ck(self._l.tl_timer_stop(self.ctx, C.byref(out)))
        return out.value

    def launch_count(self) -> int:
        out = C.c_longlong()
        self._ck(self._l.tl_launch_count(self.ctx, C.byref(out)))
        return out.value
