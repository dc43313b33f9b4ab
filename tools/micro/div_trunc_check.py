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


def main():
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
