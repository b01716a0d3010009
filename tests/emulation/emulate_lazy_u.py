"""Design check for the lazy u update of CG kernel A (TL_U_LAZY, tl_kernels_ring.cuh) and of k_cg_flush: the schedule of
launches -- A(it) applies u = (u + alpha(it-2) p(it-2)) + alpha(it-1) p(it-1) when it - first is even and >= 2, taking
p(it-2) from the ping-pong buffer it is about to overwrite; the flush applies the one or two updates that remain -- must
leave, after ANY number of iterations, exactly the bits of an update per iteration (CG.jl:93-104)."""
import numpy as np

rng = np.random.default_rng(5)


def reference(u, r, p, alphas, betas, rs):
    """CG.ur! / CG.p! per iteration: u += alpha p; p = r_new + beta p"""
    u, p = u.copy(), p.copy()
    for a, b, rn in zip(alphas, betas, rs):
        u = u + a * p
        p = b * p + rn
    return u, p


def lazy(u, r0, p_init, alphas, betas, rs, first=0):
    """The device schedule.  buf[k & 1 ^ 1] is written by A(k); hist arrays as on the device: alpha(k) is known after B(k)."""
    n = len(alphas)
    u = u.copy()
    buf = [None, None]
    pin_of = lambda it: 1 if (it & 1) else 0          # pin = (it & 1) ? p1 : p0
    buf[pin_of(first)] = p_init.copy()
    buf[1 - pin_of(first)] = rng.standard_normal(p_init.shape)     # garbage: never read before it is written
    r = r0
    for k in range(n + 1):                           # A(first + k) for k = 0..n-1 run; k = n is the flush
        it = first + k
        pin, pout = buf[pin_of(it)], buf[1 - pin_of(it)]
        if k == n:                                   # k_cg_flush<TL_U_LAZY>
            if k == 0:
                return u, p_init.copy()
            two = (k & 1) == 0
            un = u
            if two:
                un = un + alphas[k - 2] * pout       # p(it-2): the other buffer
            un = un + alphas[k - 1] * pin
            return un, betas[k - 1] * pin + r
        if k == 0:                                   # first: p copied through, nothing pending
            pnew = pin.copy()
        else:
            pnew = betas[k - 1] * pin + r            # p(it) = beta p(it-1) + r
            if k >= 2 and (k & 1) == 0:
                u = (u + alphas[k - 2] * pout) + alphas[k - 1] * pin
        buf[1 - pin_of(it)] = pnew                   # overwrites p(it-2)
        r = rs[k]                                    # B(it): r advanced, alpha(it) now known
    raise AssertionError


def run(n, first=0, shape=(7, 5)):
    u, r0, p = (rng.standard_normal(shape) for _ in range(3))
    alphas, betas = rng.uniform(0.1, 2.0, n), rng.uniform(0.1, 2.0, n)
    rs = [rng.standard_normal(shape) for _ in range(n)]
    ur, pr = reference(u, r0, p, alphas, betas, rs)
    ul, pl = lazy(u, r0, p, alphas, betas, rs, first)
    return np.array_equal(ur, ul) and np.array_equal(pr, pl)


if __name__ == "__main__":
    for first in (0, 1, 30, 31):
        for n in range(0, 12):
            assert run(n, first), (n, first)
    print("lazy u schedule OK")
