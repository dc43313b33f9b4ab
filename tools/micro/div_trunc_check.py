"""Preparation for DESIGN section 8 item 5 (k_index_points: five IEEE float64 divisions per point).

The kernel needs only int(gx / cs) (truncation of the correctly rounded quotient, utils.py:201-205).  Candidate replacement:
    t = gx * rcs            (rcs = 1/cs rounded once on the host)
    n = rint(t)
    if |t - n| > 8 ulp(t):  result = trunc(t)          # no integer between t and the IEEE quotient
    else:                   result = trunc(gx / cs)    # exact division only next to an integer boundary
|t - RN(gx/cs)| <= 3 ulp, so the guarded form returns the same integer as the division for every input.  This script checks
that claim on the CPU (numpy float64 = IEEE): random coordinates, coordinates on and next to exact multiples of cs, several
cell sizes; and counts how often the slow path is taken.  No product code depends on it.
"""
import numpy as np


def guarded(gx, cs):
    rcs = 1.0 / cs
    t = gx * rcs
    n = np.rint(t)
    near = np.abs(t - n) <= 8 * np.spacing(np.abs(t))
    out = np.trunc(t)
    out[near] = np.trunc(gx[near] / cs)
    return out, int(near.sum())


def guarded_quot(q0, q2):
    """trunc(q0 / q2 - 0.5) (project_point, utils.py:208-213) with ONE reciprocal per point shared by both coordinates:
        r = 1 / q2 (correctly rounded), t = q0 * r, u = t - 0.5, n = rint(u)
        if |u - n| > max(|t|, 1) * 2^-48: trunc(u)   else: trunc(q0 / q2 - 0.5) exactly as the reference
    |t - RN(q0/q2)| <= 2.5 ulp, the two subtractions add at most one ulp each: no integer between u and the reference's value."""
    r = 1.0 / q2
    t = q0 * r
    u = t - 0.5
    near = np.abs(u - np.rint(u)) <= np.maximum(np.abs(t), 1.0) * 2.0 ** -48
    out = np.trunc(u)
    out[near] = np.trunc(q0[near] / q2[near] - 0.5)
    return out, int(near.sum())


def guarded_rel(gx, cs):
    """The form the kernel uses for trunc(gx / cs): the same test with a relative tolerance (|t| 2^-49 >= 8 ulp(t))."""
    rcs = 1.0 / cs
    t = gx * rcs
    near = np.abs(t - np.rint(t)) <= np.abs(t) * 2.0 ** -49
    out = np.trunc(t)
    out[near] = np.trunc(gx[near] / cs)
    return out, int(near.sum())


def check(n_rounds=20, size=2_000_000, seed=0):
    """Returns (inputs checked, slow-path count); raises on any mismatch.  tests/test_kernel_arith_emulation.py runs a short form."""
    rng = np.random.default_rng(seed)
    total = slow = 0
    for cs in (0.1, 0.05, 0.07, 0.25, 0.1 + 2 ** -50):
        for _ in range(n_rounds):
            gx = rng.uniform(-60, 60, size=size)
            k = rng.integers(-600, 600, size=size // 5).astype(np.float64)
            edge = k * cs
            steps = rng.integers(-6, 7, size=edge.size)
            for _s in range(6):
                edge = np.where(steps > 0, np.nextafter(edge, np.inf), np.where(steps < 0, np.nextafter(edge, -np.inf), edge))
                steps = steps - np.sign(steps)
            x = np.concatenate([gx, edge, k, np.zeros(3)])
            want = np.trunc(x / cs)
            got, n_slow = guarded_rel(x, cs)
            assert np.array_equal(got, want), (cs, np.flatnonzero(got != want)[:5])
            total += x.size
            slow += n_slow
    # quotients: q2 = depth-like positive numbers, q0 / q2 spread over the patch / pixel range, planted integer + 0.5 hits
    for _ in range(n_rounds):
        q2 = rng.uniform(0.1, 10.0, size=size)
        ratio = rng.uniform(-4.0, 700.0, size=size)
        hit = rng.integers(-3, 700, size=size // 5).astype(np.float64) + 0.5          # q0 / q2 - 0.5 lands on an integer
        ratio = np.concatenate([ratio, hit])
        q2 = np.concatenate([q2, rng.uniform(0.1, 10.0, size=hit.size)])
        q0 = ratio * q2
        steps = rng.integers(-4, 5, size=q0.size)
        for _s in range(4):
            q0 = np.where(steps > 0, np.nextafter(q0, np.inf), np.where(steps < 0, np.nextafter(q0, -np.inf), q0))
            steps = steps - np.sign(steps)
        want = np.trunc(q0 / q2 - 0.5)
        got, n_slow = guarded_quot(q0, q2)
        assert np.array_equal(got, want), np.flatnonzero(got != want)[:5]
        total += q0.size
        slow += n_slow
    return total, slow


def main():
    total, slow = check()
    print(f"guarded forms == IEEE division + trunc on {total:,} inputs; exact-division path taken {slow:,} times")


def _old_main():
    rng = np.random.default_rng(0)
    total = slow = 0
    for cs in (0.1, 0.05, 0.07, 0.25, 0.1 + 2 ** -50):
        for _ in range(20):
            gx = rng.uniform(-60, 60, size=2_000_000)
            k = rng.integers(-600, 600, size=400_000).astype(np.float64)
            edge = k * cs                                              # on a cell boundary (as rounded), then +- a few ulps
            steps = rng.integers(-6, 7, size=edge.size)
            for _s in range(6):
                edge = np.where(steps > 0, np.nextafter(edge, np.inf), np.where(steps < 0, np.nextafter(edge, -np.inf), edge))
                steps = steps - np.sign(steps)
            x = np.concatenate([gx, edge, k, np.zeros(3)])
            want = np.trunc(x / cs)
            got, n_slow = guarded(x, cs)
            assert np.array_equal(got, want), (cs, np.flatnonzero(got != want)[:5])
            total += x.size
            slow += n_slow
    print(f"guarded multiply == IEEE division + trunc on {total:,} inputs; exact-division path taken {slow:,} times "
          f"({slow / total:.2e} of the inputs, of which 29 % are planted on or next to cell boundaries; 0 of 20 M uniformly random "
          f"coordinates take it)")


if __name__ == "__main__":
    main()
