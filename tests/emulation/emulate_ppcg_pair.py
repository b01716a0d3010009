"""NumPy emulation of k_ppcg_pair_ring's carry logic vs two single inner steps (bit for bit)."""
import numpy as np
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emulate_pair import stencil_full, stencil2  # noqa: E402
rng = np.random.default_rng(7)

def single_inner(sd, r, u, kx, ky, nx, ny, hd, al, be):
    w = stencil_full(sd, kx, ky, nx, ny, hd)
    I = (slice(hd, hd + ny), slice(hd, hd + nx))
    rn = r[I] - w
    un = u[I] + sd[I]
    sn = al * sd[I] + be * rn
    r2, u2, s2 = r.copy(), u.copy(), sd.copy()
    r2[I] = rn; u2[I] = un; s2[I] = sn
    return s2, r2, u2

def pair_kernel(sd, r, u, kx, ky, nx, ny, hd, aA, bA, aB, bB, rpc, OWN=60):
    def ld2(f, j, i0, ok):
        out = np.zeros((32, 2))
        for l in range(32):
            if ok[l]:
                out[l, 0] = f[j + hd, i0[l] + hd]; out[l, 1] = f[j + hd, i0[l] + 1 + hd]
        return out
    sout = np.full_like(sd, np.nan); rout = np.full_like(r, np.nan); uo = u.copy()
    lane = np.arange(32)
    for q in range(-(-ny // rpc)):
        for s in range(-(-nx // OWN)):
            j0 = q * rpc; j1 = min(ny, j0 + rpc)
            own_lo = s * OWN; own_hi = min(nx, own_lo + OWN)
            i0 = own_lo - 2 + 2 * lane
            ok = i0 <= nx
            own_a = (i0 >= own_lo) & (i0 < own_hi); own_b = (i0 + 1 >= own_lo) & (i0 + 1 < own_hi) & own_a
            ja_lo = 0 if j0 == 0 else j0 - 1
            ja_hi = ny - 1 if j1 == ny else j1
            jm = 0 if ja_lo == 0 else ja_lo - 1
            Sm = ld2(sd, jm, i0, ok); Sc = ld2(sd, ja_lo, i0, ok); kyc = ld2(ky, ja_lo, i0, ok)
            Z = np.zeros((32, 2)); Am, Ac, rAc, uAc, kxc, kyB = Z, Z, Z, Z, Z, Z
            def step_b(j, An, kyn):
                Bm = Ac if j == 0 else Am
                w = stencil2(nx, i0, lane, Bm, Ac, An, kxc, kyB, kyn)
                rn = rAc - w; un = uAc + Ac; sn = aB * Ac + bB * rn
                for l in range(32):
                    for c, own in ((0, own_a), (1, own_b)):
                        if own[l]:
                            rout[j + hd, i0[l] + c + hd] = rn[l, c]; uo[j + hd, i0[l] + c + hd] = un[l, c]; sout[j + hd, i0[l] + c + hd] = sn[l, c]
            for jj in range(ja_lo, ja_hi + 1):
                jn = ny - 1 if jj + 1 >= ny else jj + 1
                Sn = ld2(sd, jn, i0, ok); kyn_ = ld2(ky, jj + 1, i0, ok); kxv = ld2(kx, jj, i0, ok)
                a = ld2(r, jj, i0, ok); b = ld2(u, jj, i0, ok)
                w = stencil2(nx, i0, lane, Sm, Sc, Sn, kxv, kyc, kyn_)
                rA = a - w; uA = b + Sc; sA = aA * Sc + bA * rA
                if jj - 1 >= j0:
                    step_b(jj - 1, sA, kyc)
                Am = Ac; Ac = sA; rAc = rA; uAc = uA; kxc = kxv; kyB = kyc
                Sm = Sc; Sc = Sn; kyc = kyn_
            if ja_hi == j1 - 1:
                step_b(j1 - 1, Ac, kyc)
    return sout, rout, uo

CASES = [(7, 5, 2), (61, 6, 3), (121, 7, 2), (1, 6, 3), (130, 1, 4), (59, 10, 32), (120, 4, 1)]


def run(nx, ny, rpc):
    hd = 2
    shape = (ny + 2 * hd, nx + 2 * hd + 70)
    f = lambda: rng.standard_normal(shape)
    sd, r, u = f(), f(), f(); kx, ky = np.abs(f()), np.abs(f())
    aA, bA, aB, bB = 0.37, 0.011, 0.41, 0.013
    s1, r1, u1 = single_inner(sd, r, u, kx, ky, nx, ny, hd, aA, bA)
    s2, r2, u2 = single_inner(s1, r1, u1, kx, ky, nx, ny, hd, aB, bB)
    so, ro, uo = pair_kernel(sd, r, u, kx, ky, nx, ny, hd, aA, bA, aB, bB, rpc)
    I = (slice(hd, hd + ny), slice(hd, hd + nx))
    ok = np.array_equal(so[I], s2[I]) and np.array_equal(ro[I], r2[I]) and np.array_equal(uo[I], u2[I])
    print(f"{nx}x{ny} rows/chunk {rpc}: bit-identical = {ok}")
    assert ok


if __name__ == "__main__":
    for nx, ny, rpc in CASES:
        run(nx, ny, rpc)
    print("ppcg pair emulation OK")
