"""Design check for the TILED two-inner-steps-per-pass PPCG kernel (k_ppcg_pair_ring<.., TILED>):
px x py tiles with halos of depth 2; before a pass the tile-internal halos hold the neighbours' sd two cells deep (corner
blocks included), r one cell deep, kx / ky two deep, and NOTHING of u (u is only ever updated at owned cells).  Every tile
runs the same window kernel as the single-tile emulation (emulate_ppcg_pair.py), clamping on physical sides only.  The
assembled result must equal two single-chunk inner steps (PPCG.jl:78-83) bit for bit; unfilled halo cells hold NaN, so
any use of a cell outside the stated depths poisons the output."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emulate_pair_tiled import stencil2_phys, split  # noqa: E402
from emulate_ppcg_pair import single_inner  # noqa: E402

rng = np.random.default_rng(11)
HD = 2


def ppcg_pair_tile(sd, r, u, kx, ky, nx, ny, phys, aA, bA, aB, bB, rpc, OWN=60):
    """One tile; arrays are (ny + 2 HD, nx + 2 HD + pad), interior origin (HD, HD); phys = (L, R, B, T)."""
    physL, physR, physB, physT = phys
    W = sd.shape[1]

    def ld2(f, j, i0, ok):
        out = np.zeros((32, 2))
        for l in range(32):
            if ok[l] and i0[l] + HD >= 0 and i0[l] + 1 + HD < W:
                out[l, 0] = f[j + HD, i0[l] + HD]; out[l, 1] = f[j + HD, i0[l] + 1 + HD]
        return out
    sout = np.full_like(sd, np.nan); rout = np.full_like(r, np.nan); uout = np.full_like(u, np.nan)
    lane = np.arange(32)
    for q in range(-(-ny // rpc)):
        for s in range(-(-nx // OWN)):
            j0 = q * rpc; j1 = min(ny, j0 + rpc)
            own_lo = s * OWN; own_hi = min(nx, own_lo + OWN)
            i0 = own_lo - 2 + 2 * lane
            ok = i0 <= nx + (0 if physR else 1)
            own_a = (i0 >= own_lo) & (i0 < own_hi); own_b = (i0 + 1 >= own_lo) & (i0 + 1 < own_hi) & own_a
            ja_lo = 0 if (j0 == 0 and physB) else j0 - 1
            ja_hi = ny - 1 if (j1 == ny and physT) else j1
            jm = 0 if (ja_lo == 0 and physB) else ja_lo - 1
            Sm = ld2(sd, jm, i0, ok); Sc = ld2(sd, ja_lo, i0, ok); kyc = ld2(ky, ja_lo, i0, ok)
            Z = np.zeros((32, 2)); Am, Ac, rAc, uAc, kxc, kyB = Z, Z, Z, Z, Z, Z

            def step_b(j, An, kyn):
                Bm = Ac if (j == 0 and physB) else Am
                w = stencil2_phys(nx, i0, lane, phys, Bm, Ac, An, kxc, kyB, kyn)
                rn = rAc - w; un = uAc + Ac; sn = aB * Ac + bB * rn
                for l in range(32):
                    for c, own in ((0, own_a), (1, own_b)):
                        if own[l]:
                            rout[j + HD, i0[l] + c + HD] = rn[l, c]; uout[j + HD, i0[l] + c + HD] = un[l, c]
                            sout[j + HD, i0[l] + c + HD] = sn[l, c]
            for jj in range(ja_lo, ja_hi + 1):
                jn = ny - 1 if (jj + 1 >= ny and physT) else jj + 1
                Sn = ld2(sd, jn, i0, ok); kyn_ = ld2(ky, jj + 1, i0, ok); kxv = ld2(kx, jj, i0, ok)
                a = ld2(r, jj, i0, ok); b = ld2(u, jj, i0, ok)
                w = stencil2_phys(nx, i0, lane, phys, Sm, Sc, Sn, kxv, kyc, kyn_)
                rA = a - w; uA = b + Sc; sA = aA * Sc + bA * rA
                if jj - 1 >= j0:
                    step_b(jj - 1, sA, kyc)
                Am = Ac; Ac = sA; rAc = rA; uAc = uA; kxc = kxv; kyB = kyc
                Sm = Sc; Sc = Sn; kyc = kyn_
            if ja_hi == j1 - 1:
                step_b(j1 - 1, Ac, kyc)
    return sout, rout, uout


def run(NX, NY, px, py, rpc, ds=2, dr=1, du=0, dk=2):
    """ds / dr / du / dk: halo depth filled for sd / r / u / (kx, ky) on tile-internal sides."""
    PAD = 70
    G = lambda: rng.standard_normal((NY + 2 * HD, NX + 2 * HD + PAD))
    sd, r, u = G(), G(), G(); kx, ky = np.abs(G()), np.abs(G())
    aA, bA, aB, bB = 0.37, 0.011, 0.41, 0.013
    s1, r1, u1 = single_inner(sd, r, u, kx, ky, NX, NY, HD, aA, bA)
    s2, r2, u2 = single_inner(s1, r1, u1, kx, ky, NX, NY, HD, aB, bB)
    S = np.full_like(sd, np.nan); R = np.full_like(r, np.nan); U = np.full_like(u, np.nan)
    for cy in range(py):
        for cx in range(px):
            x0, nx = split(NX, px, cx); y0, ny = split(NY, py, cy)
            phys = (cx == 0, cx == px - 1, cy == 0, cy == py - 1)

            def window(depth):
                lo_x = -depth if not phys[0] else 0; hi_x = nx + (depth if not phys[1] else 0)
                lo_y = -depth if not phys[2] else 0; hi_y = ny + (depth if not phys[3] else 0)
                return lo_x, hi_x, lo_y, hi_y

            def tile_of(f, depth):
                t = np.full((ny + 2 * HD, nx + 2 * HD + PAD), np.nan)
                lo_x, hi_x, lo_y, hi_y = window(depth)
                t[HD + lo_y:HD + hi_y, HD + lo_x:HD + hi_x] = f[HD + y0 + lo_y:HD + y0 + hi_y, HD + x0 + lo_x:HD + x0 + hi_x]
                g = rng.standard_normal(t.shape)          # physical halos: garbage, never read past the clamp
                if phys[0]: t[:, :HD] = g[:, :HD]
                if phys[1]: t[:, HD + nx:] = np.where(np.isnan(t[:, HD + nx:]), g[:, HD + nx:], t[:, HD + nx:])
                if phys[2]: t[:HD, :] = np.where(np.isnan(t[:HD, :]), g[:HD, :], t[:HD, :])
                if phys[3]: t[HD + ny:, :] = np.where(np.isnan(t[HD + ny:, :]), g[HD + ny:, :], t[HD + ny:, :])
                return t
            ts, tr, tu = tile_of(sd, ds), tile_of(r, dr), tile_of(u, du)
            tkx, tky = tile_of(kx, dk), tile_of(ky, dk)
            lo_x, hi_x, lo_y, hi_y = window(dk)      # the coefficient row / column ON a physical top / right side is real data
            if phys[3]:
                tky[HD + ny, HD + lo_x:HD + hi_x] = ky[HD + y0 + ny, HD + x0 + lo_x:HD + x0 + hi_x]
            if phys[1]:
                tkx[HD + lo_y:HD + hi_y, HD + nx] = kx[HD + y0 + lo_y:HD + y0 + hi_y, HD + x0 + nx]
            so, ro, uo = ppcg_pair_tile(ts, tr, tu, tkx, tky, nx, ny, phys, aA, bA, aB, bB, rpc)
            T = (slice(HD + y0, HD + y0 + ny), slice(HD + x0, HD + x0 + nx)); t = (slice(HD, HD + ny), slice(HD, HD + nx))
            S[T] = so[t]; R[T] = ro[t]; U[T] = uo[t]
    I = (slice(HD, HD + NY), slice(HD, HD + NX))
    return np.array_equal(S[I], s2[I]) and np.array_equal(R[I], r2[I]) and np.array_equal(U[I], u2[I])


CASES = [(70, 9, 2, 1, 4), (20, 30, 1, 2, 4), (130, 12, 2, 2, 3), (61, 10, 3, 3, 2), (9, 9, 3, 3, 32)]

if __name__ == "__main__":
    for c in CASES:
        print(c, "sufficient depths (sd 2, r 1, u 0, kx/ky 2):", run(*c))
        assert run(*c)
    print("sd only 1 deep:", run(130, 12, 2, 2, 3, ds=1), " r 0 deep:", run(130, 12, 2, 2, 3, dr=0), " kx/ky 1 deep:", run(130, 12, 2, 2, 3, dk=1))
