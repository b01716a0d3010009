"""Design check for the TILED two-iterations-per-pass Chebyshev kernel: px x py tiles,
each with a halo of depth 2; before a pass the tile-internal halos hold the neighbours' u two cells deep
(corner blocks included), p and u0 one cell deep and kx, ky two deep; every tile then runs the SAME window
kernel as the single-tile emulation (emulate_pair.py), clamping only on physical sides.  The assembled
result must equal two single-chunk iterations bit for bit, and the halo depths above must be sufficient
(every halo cell that was not filled holds NaN: any use would poison the output)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emulate_pair import single_step, shfl_up, shfl_down  # noqa: E402

rng = np.random.default_rng(3)
HD = 2


def stencil2_phys(nx, i0, lane, phys, Xm, Xc, Xn, kxv, kyc, kyn):
    physL, physR = phys[0], phys[1]
    xl = shfl_up(Xc[:, 1]); xr = shfl_down(Xc[:, 0]); kxr = shfl_down(kxv[:, 0])
    xl = np.where(lane == 0, 0.0, xl); xr = np.where(lane == 31, 0.0, xr); kxr = np.where(lane == 31, 0.0, kxr)
    La = np.where(physL & (i0 == 0), Xc[:, 0], xl)
    Ra = np.where(physR & (i0 == nx - 1), Xc[:, 0], Xc[:, 1])
    Lb = Xc[:, 0]
    Rb = np.where(physR & (i0 + 1 == nx - 1), Xc[:, 1], xr)
    wa = ((((1.0 + kxv[:, 1]) + kxv[:, 0]) + kyn[:, 0]) + kyc[:, 0]) * Xc[:, 0] - (kxv[:, 1] * Ra + kxv[:, 0] * La) - (kyn[:, 0] * Xn[:, 0] + kyc[:, 0] * Xm[:, 0])
    wb = ((((1.0 + kxr) + kxv[:, 1]) + kyn[:, 1]) + kyc[:, 1]) * Xc[:, 1] - (kxr * Rb + kxv[:, 1] * Lb) - (kyn[:, 1] * Xn[:, 1] + kyc[:, 1] * Xm[:, 1])
    return np.stack([wa, wb], 1)


def pair_kernel_tile(u, p, u0, kx, ky, nx, ny, phys, aA, bA, aB, bB, rpc, OWN=60):
    """One tile; arrays are (ny + 2 HD, nx + 2 HD + pad) with interior origin (HD, HD); phys = (L, R, B, T)."""
    physL, physR, physB, physT = phys
    W = u.shape[1]

    def ld2(f, j, i0, ok):
        out = np.zeros((32, 2))
        for l in range(32):
            if ok[l] and i0[l] + HD >= 0 and i0[l] + 1 + HD < W:
                out[l, 0] = f[j + HD, i0[l] + HD]; out[l, 1] = f[j + HD, i0[l] + 1 + HD]
        return out
    uout = np.full_like(u, np.nan); pout = np.full_like(p, np.nan)
    lane = np.arange(32)
    for q in range(-(-ny // rpc)):
        for s in range(-(-nx // OWN)):
            j0 = q * rpc; j1 = min(ny, j0 + rpc)
            own_lo = s * OWN; own_hi = min(nx, own_lo + OWN)
            i0 = own_lo - 2 + 2 * lane
            # a physical right side never needs more than the pair (nx, nx+1); a tile-internal one needs u(nx+1) for the
            # redundant uA(nx), which for odd nx sits in the pair (nx+1, nx+2)   [found by this emulation]
            ok = i0 <= nx + (0 if physR else 1)
            own_a = (i0 >= own_lo) & (i0 < own_hi); own_b = (i0 + 1 >= own_lo) & (i0 + 1 < own_hi) & own_a
            ja_lo = 0 if (j0 == 0 and physB) else j0 - 1
            ja_hi = ny - 1 if (j1 == ny and physT) else j1
            jm = 0 if (ja_lo == 0 and physB) else ja_lo - 1
            Um = ld2(u, jm, i0, ok); Uc = ld2(u, ja_lo, i0, ok); kyc = ld2(ky, ja_lo, i0, ok)
            Z = np.zeros((32, 2)); Am, Ac, pAc, u0c, kxc, kyB = Z, Z, Z, Z, Z, Z

            def step_b(j, An, kyn):
                Bm = Ac if (j == 0 and physB) else Am
                w = stencil2_phys(nx, i0, lane, phys, Bm, Ac, An, kxc, kyB, kyn)
                r = u0c - w
                pn = aB * pAc + bB * r
                un = Ac + pn
                for l in range(32):
                    for c, own in ((0, own_a), (1, own_b)):
                        if own[l]:
                            pout[j + HD, i0[l] + c + HD] = pn[l, c]; uout[j + HD, i0[l] + c + HD] = un[l, c]
            for jj in range(ja_lo, ja_hi + 1):
                jn = ny - 1 if (jj + 1 >= ny and physT) else jj + 1
                Un = ld2(u, jn, i0, ok); kyn_ = ld2(ky, jj + 1, i0, ok); kxv = ld2(kx, jj, i0, ok)
                a = ld2(u0, jj, i0, ok); b = ld2(p, jj, i0, ok)
                w = stencil2_phys(nx, i0, lane, phys, Um, Uc, Un, kxv, kyc, kyn_)
                r = a - w
                pA = aA * b + bA * r
                uA = Uc + pA
                if jj - 1 >= j0:
                    step_b(jj - 1, uA, kyc)
                Am = Ac; Ac = uA; pAc = pA; u0c = a; kxc = kxv; kyB = kyc
                Um = Uc; Uc = Un; kyc = kyn_
            if ja_hi == j1 - 1:
                step_b(j1 - 1, Ac, kyc)
    return uout, pout


def split(n, parts, i):
    base, rem = divmod(n, parts)
    lo = i * base + min(i, rem)
    return lo, base + (1 if i < rem else 0)


def run(NX, NY, px, py, rpc, du=2, dp=1, dk=2):
    """du / dp / dk: halo depth filled for u / (p, u0) / (kx, ky) on tile-internal sides."""
    PAD = 70
    G = lambda: rng.standard_normal((NY + 2 * HD, NX + 2 * HD + PAD))
    u, p, u0 = G(), G(), G(); kx, ky = np.abs(G()), np.abs(G())
    aA, bA, aB, bB = 0.37, 0.011, 0.41, 0.013
    u1, p1, _, _ = single_step(u, p, u0, kx, ky, NX, NY, HD, aA, bA)
    u2, p2, _, _ = single_step(u1, p1, u0, kx, ky, NX, NY, HD, aB, bB)
    U = np.full_like(u, np.nan); P = np.full_like(p, np.nan)
    for cy in range(py):
        for cx in range(px):
            x0, nx = split(NX, px, cx); y0, ny = split(NY, py, cy)
            phys = (cx == 0, cx == px - 1, cy == 0, cy == py - 1)

            def tile_of(f, depth, physical_garbage=True):
                t = np.full((ny + 2 * HD, nx + 2 * HD + PAD), np.nan)
                # interior
                t[HD:HD + ny, HD:HD + nx] = f[HD + y0:HD + y0 + ny, HD + x0:HD + x0 + nx]
                # tile-internal halos, `depth` deep, corner blocks included; physical sides: garbage (never read past the clamp)
                lo_x = -depth if not phys[0] else 0; hi_x = nx + (depth if not phys[1] else 0)
                lo_y = -depth if not phys[2] else 0; hi_y = ny + (depth if not phys[3] else 0)
                t[HD + lo_y:HD + hi_y, HD + lo_x:HD + hi_x] = f[HD + y0 + lo_y:HD + y0 + hi_y, HD + x0 + lo_x:HD + x0 + hi_x]
                if physical_garbage:
                    g = rng.standard_normal(t.shape)
                    if phys[0]: t[:, :HD] = g[:, :HD]
                    if phys[1]: t[:, HD + nx:] = np.where(np.isnan(t[:, HD + nx:]), g[:, HD + nx:], t[:, HD + nx:])
                    if phys[2]: t[:HD, :] = np.where(np.isnan(t[:HD, :]), g[:HD, :], t[:HD, :])
                    if phys[3]: t[HD + ny:, :] = np.where(np.isnan(t[HD + ny:, :]), g[HD + ny:, :], t[HD + ny:, :])
                return t
            tu, tp, tu0 = tile_of(u, du), tile_of(p, dp), tile_of(u0, dp)
            tkx, tky = tile_of(kx, dk), tile_of(ky, dk)
            # ky(row ny) / kx(col nx) of a PHYSICAL top/right side are read by the stencil as in the single-chunk code, and
            # the redundant uA cells read them in the tile-internal halo columns / rows too: the wide pull of kx, ky must
            # cover the neighbours' physical halo row / column   [found by this emulation]
            lo_x = -dk if not phys[0] else 0; hi_x = nx + (dk if not phys[1] else 0)
            lo_y = -dk if not phys[2] else 0; hi_y = ny + (dk if not phys[3] else 0)
            if phys[3]:
                tky[HD + ny, HD + lo_x:HD + hi_x] = ky[HD + y0 + ny, HD + x0 + lo_x:HD + x0 + hi_x]
            if phys[1]:
                tkx[HD + lo_y:HD + hi_y, HD + nx] = kx[HD + y0 + lo_y:HD + y0 + hi_y, HD + x0 + nx]
            uo, po = pair_kernel_tile(tu, tp, tu0, tkx, tky, nx, ny, phys, aA, bA, aB, bB, rpc)
            U[HD + y0:HD + y0 + ny, HD + x0:HD + x0 + nx] = uo[HD:HD + ny, HD:HD + nx]
            P[HD + y0:HD + y0 + ny, HD + x0:HD + x0 + nx] = po[HD:HD + ny, HD:HD + nx]
    I = (slice(HD, HD + NY), slice(HD, HD + NX))
    return np.array_equal(U[I], u2[I]) and np.array_equal(P[I], p2[I])


CASES = [(70, 9, 2, 1, 4), (20, 30, 1, 2, 4), (130, 12, 2, 2, 3), (61, 10, 3, 3, 2), (9, 9, 3, 3, 32)]

if __name__ == "__main__":
    for c in CASES:
        print(c, "sufficient depths (u 2, p/u0 1, kx/ky 2):", run(*c))
        assert run(*c)
    # necessity: one cell less of u, or no p halo, must NOT work (NaN poisons the result) on a decomposed mesh
    print("u only 1 deep:", run(130, 12, 2, 2, 3, du=1), " p/u0 0 deep:", run(130, 12, 2, 2, 3, dp=0), " kx/ky 1 deep:", run(130, 12, 2, 2, 3, dk=1))
