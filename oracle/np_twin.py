"""NumPy twin of the oracle -- an independent, vectorised restatement of the same intended
algorithm (src/solvers/{CG,Cheby,PPCG}.jl + src/kernels.jl with SURVEY.md Appendix A).

TEST INFRASTRUCTURE ONLY.  Its job is to pin oracle/tealeaf_oracle.c: the two were written
separately (scalar C loops vs whole-array NumPy expressions, QL iteration vs LAPACK for the
eigenvalues), so agreement between them is the strongest check available while the reference
itself cannot be executed (PARITY UNPINNED, see the C header).  Element-wise results (w = A u,
kx, ky, ...) must agree bit for bit -- same expression order; sums agree to rounding (NumPy
sums pairwise, the C oracle serially).
"""
from __future__ import annotations

import math

import numpy as np

ERROR_START = 1e10
ERROR_SWITCH_MAX = 1.0
CGEIGENITERS = 20


class Twin:
    def __init__(self, density, energy, hd, coef=1):
        self.hd = hd
        self.density = np.array(density, dtype=np.float64, order="F")
        self.energy = np.array(energy, dtype=np.float64, order="F")
        self.x, self.y = self.density.shape
        z = lambda: np.zeros((self.x, self.y), order="F")
        self.u, self.u0, self.p, self.r, self.w, self.kx, self.ky, self.sd = (z() for _ in range(8))
        self.coef = coef
        self.I = (slice(hd, self.x - hd), slice(hd, self.y - hd))
        self.cgalpha, self.cgbeta = [], []

    # -- src/kernels.jl:191-210 (A#2) --
    def halo(self, a, depth=1):
        hd, x, y = self.hd, self.x, self.y
        for d in range(1, depth + 1):
            a[hd - d, hd:y - hd] = a[hd + d - 1, hd:y - hd]
            a[x - hd + d - 1, hd:y - hd] = a[x - hd - d, hd:y - hd]
        for d in range(1, depth + 1):
            a[hd:x - hd, y - hd + d - 1] = a[hd:x - hd, y - hd - d]
            a[hd:x - hd, hd - d] = a[hd:x - hd, hd + d - 1]

    # -- src/kernels.jl:178-184 as one expression (A#1) --
    def smvp(self, a):
        hd, x, y = self.hd, self.x, self.y
        c = (slice(hd, x - hd), slice(hd, y - hd))
        xr = (slice(hd + 1, x - hd + 1), slice(hd, y - hd))
        xl = (slice(hd - 1, x - hd - 1), slice(hd, y - hd))
        yu = (slice(hd, x - hd), slice(hd + 1, y - hd + 1))
        yd = (slice(hd, x - hd), slice(hd - 1, y - hd - 1))
        kx, ky = self.kx, self.ky
        consum = (((1.0 + kx[xr]) + kx[c]) + ky[yu]) + ky[c]
        return consum * a[c] - (kx[xr] * a[xr] + kx[c] * a[xl]) - (ky[yu] * a[yu] + ky[c] * a[yd])

    # -- src/solvers/CG.jl:47-79 --
    def cg_init(self, rx, ry):
        hd, x, y = self.hd, self.x, self.y
        self.u = self.energy * self.density
        self.p[...] = 0.0
        self.r[...] = 0.0
        d = self.density[1:-1, 1:-1]
        self.w[1:-1, 1:-1] = d if self.coef == 1 else 1.0 / d
        w = self.w
        K = (slice(hd, x - 1), slice(hd, y - 1))
        wl = w[hd - 1:x - 2, hd:y - 1]
        wd = w[hd:x - 1, hd - 1:y - 2]
        self.kx[K] = rx * (wl + w[K]) / (2.0 * wl * w[K])
        self.ky[K] = ry * (wd + w[K]) / (2.0 * wd * w[K])
        I = self.I
        self.w[I] = self.smvp(self.u)
        self.r[I] = self.u[I] - self.w[I]
        self.p[I] = self.r[I]
        return float(np.sum(self.r[I] ** 2))

    def cg_mainstep(self, rro):
        I = self.I
        self.w[I] = self.smvp(self.p)
        pw = float(np.sum(self.w[I] * self.p[I]))
        alpha = rro / pw
        self.cgalpha.append(alpha)
        self.u[I] += alpha * self.p[I]
        self.r[I] -= alpha * self.w[I]
        rrn = float(np.sum(self.r[I] ** 2))
        beta = rrn / rro
        self.cgbeta.append(beta)
        self.p[I] = beta * self.p[I] + self.r[I]
        return rrn

    def preamble(self, rx, ry):
        rro = self.cg_init(rx, ry)
        self.halo(self.u); self.halo(self.p)
        self.u0[self.I] = self.u[self.I]
        self.cgalpha, self.cgbeta = [], []
        return rro

    # -- src/solvers/CG.jl:7-29 --
    def cg_solve(self, rx, ry, eps, maxiters):
        rro = self.preamble(rx, ry)
        error, iters = ERROR_START, 0
        for tt in range(1, maxiters + 1):
            iters = tt
            rro = error = self.cg_mainstep(rro)
            self.halo(self.u); self.halo(self.p)
            if math.sqrt(abs(error)) < eps:
                break
        return {"iters": iters, "error": error}

    # -- src/kernels.jl:19-51 (A#9-11) --
    def eigenvalues(self, n):
        a, b = np.array(self.cgalpha[:n]), np.array(self.cgbeta[:n])
        diag = 1.0 / a
        diag[1:] += b[:-1] / a[:-1]
        off = np.sqrt(b[:-1]) / a[:-1]
        ev = np.linalg.eigvalsh(np.diag(diag) + np.diag(off, 1) + np.diag(off, -1))
        return ev[0] * 0.95, ev[-1] * 1.05

    @staticmethod
    def coef_arrays(eigmin, eigmax, n):
        theta = (eigmax + eigmin) / 2
        delta = (eigmax - eigmin) / 2
        sigma = theta / delta
        rho_old = 1 / sigma
        al, be = [], []
        for _ in range(n):
            rho_new = 1 / (2 * sigma - rho_old)
            al.append(rho_new * rho_old); be.append(2 * rho_new / delta)
            rho_old = rho_new
        return theta, al, be

    @staticmethod
    def switch(errorswitch, error, epslim, tt, presteps):
        return (error < epslim and tt > CGEIGENITERS) if errorswitch else (tt > presteps and error < ERROR_SWITCH_MAX)

    # -- src/solvers/Cheby.jl:10-61 (A#12-14) --
    def cheby_solve(self, rx, ry, eps, maxiters, presteps=30, epslim=1e-5, errorswitch=False):
        I = self.I
        rro = self.preamble(rx, ry)
        error, estiter, chebyiters, cgiters = ERROR_START, 0, 0, 0
        tt = 0
        for tt in range(1, maxiters + 1):
            if not (chebyiters != 0 or self.switch(errorswitch, error, epslim, tt, presteps)):
                rro = error = self.cg_mainstep(rro)
                cgiters = tt
                self.halo(self.u); self.halo(self.p)
            else:
                chebyiters += 1
                if chebyiters == 1:
                    eigmin, eigmax = self.eigenvalues(tt - 1)
                    theta, al, be = self.coef_arrays(eigmin, eigmax, maxiters - (tt - 1) + 2)
                    bb = float(np.sum(self.u0[I] ** 2))
                    self.w[I] = self.smvp(self.u)
                    self.r[I] = self.u0[I] - self.w[I]
                    self.p[I] = self.r[I] / theta
                    self.u[I] += self.p[I]
                    self.halo(self.u)
                calc2norm = chebyiters == 1 or (chebyiters >= estiter and (tt + 1) % 10 == 0)
                self.w[I] = self.smvp(self.u)
                self.r[I] = self.u0[I] - self.w[I]
                self.p[I] = al[chebyiters] * self.p[I] + be[chebyiters] * self.r[I]
                self.u[I] += self.p[I]
                if calc2norm:
                    error = float(np.sum(self.r[I] ** 2))
                if chebyiters == 1:
                    connum = eigmax / eigmin
                    it_alpha = np.finfo(float).eps * bb / (4 * error)
                    gamma = (math.sqrt(connum) - 1) / (math.sqrt(connum) + 1)
                    estiter = int(np.rint(math.log(it_alpha) / (2 * math.log(gamma))))
                self.halo(self.u)
            if abs(error) < eps:
                break
        return {"iters": tt, "cg_iters": cgiters, "cheby_iters": chebyiters, "est_iters": estiter, "error": error}

    # -- src/solvers/PPCG.jl:9-100 (A#12, #14-20) --
    def ppcg_solve(self, rx, ry, eps, maxiters, presteps=30, epslim=1e-5, errorswitch=False, inner=10):
        I = self.I
        rro = self.preamble(rx, ry)
        error, ppcgiters, cgiters = ERROR_START, 0, 0
        tt = 0
        for tt in range(1, maxiters + 1):
            if not (ppcgiters != 0 or self.switch(errorswitch, error, epslim, tt, presteps)):
                rro = error = self.cg_mainstep(rro)
                cgiters = tt
                self.halo(self.u); self.halo(self.p)
            else:
                ppcgiters += 1
                if ppcgiters == 1:
                    eigmin, eigmax = self.eigenvalues(tt - 1)
                    theta, al, be = self.coef_arrays(eigmin, eigmax, inner)
                    self.r[I] = self.u0[I] - self.smvp(self.u)
                    self.halo(self.p)
                    rro = float(np.sum(self.r[I] ** 2))
                self.w[I] = self.smvp(self.p)
                pw = float(np.sum(self.w[I] * self.p[I]))
                alpha = rro / pw
                self.u[I] += alpha * self.p[I]
                self.r[I] -= alpha * self.w[I]
                self.sd[I] = self.r[I] / theta
                for pp in range(inner):
                    self.halo(self.sd)
                    self.r[I] -= self.smvp(self.sd)
                    self.u[I] += self.sd[I]
                    self.sd[I] = al[pp] * self.sd[I] + be[pp] * self.r[I]
                rrn = float(np.sum(self.r[I] ** 2))
                beta = rrn / rro
                self.p[I] = beta * self.p[I] + self.r[I]
                rro = error = rrn
                self.halo(self.p)
            if abs(error) < eps:
                break
        return {"iters": tt, "cg_iters": cgiters, "cheby_iters": ppcgiters, "error": error}

    # -- src/solvers/Jacobi.jl:7-82 (A#21) --
    def jacobi_solve(self, rx, ry, eps, maxiters):
        hd, x, y = self.hd, self.x, self.y
        self.u = self.energy * self.density
        self.u0 = self.u.copy()
        d = self.density if self.coef == 1 else 1.0 / self.density
        dc, dl, dd = d[hd:x - 1, hd:y - 1], d[hd - 1:x - 2, hd:y - 1], d[hd:x - 1, hd - 1:y - 2]
        self.kx[hd:x - 1, hd:y - 1] = rx * (dl + dc) / (2.0 * dl * dc)
        self.ky[hd:x - 1, hd:y - 1] = ry * (dd + dc) / (2.0 * dd * dc)
        c = self.I
        xr = (slice(hd + 1, x - hd + 1), slice(hd, y - hd))
        xl = (slice(hd - 1, x - hd - 1), slice(hd, y - hd))
        yu = (slice(hd, x - hd), slice(hd + 1, y - hd + 1))
        yd = (slice(hd, x - hd), slice(hd - 1, y - hd - 1))
        kx, ky = self.kx, self.ky
        error, iters = ERROR_START, 0
        for tt in range(1, maxiters + 1):
            iters = tt
            self.r = self.u.copy()
            r = self.r
            num = (((self.u0[c] + kx[xr] * r[xr]) + kx[c] * r[xl]) + ky[yu] * r[yu]) + ky[c] * r[yd]
            den = (((1.0 + kx[c]) + kx[xr]) + ky[c]) + ky[yu]
            self.u[c] = num / den
            error = float(np.abs(self.u[c] - r[c]).sum())
            if tt % 50 == 0:
                self.halo(self.u)
                self.r[c] = self.u0[c] - self.smvp(self.u)
                error = float((self.r[c] ** 2).sum())
            self.halo(self.u)
            if abs(error) < eps:
                break
        return {"iters": iters, "error": error}

