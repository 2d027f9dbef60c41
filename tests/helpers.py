"""Shared helpers of the test suite (exact-number plumbing between liblsx outputs, the oracle and
the golden fixtures)."""
from fractions import Fraction

import numpy as np

from linalg_solver_b200.convert import limbs_to_ints, reduce_pq


def pq_grid_from_num(num_rows, den, rank=None, extra_den=1):
    """[[p, q], ...] row-major from integer numerators over one denominator."""
    out = []
    for i, row in enumerate(num_rows):
        d = den if (rank is None or i < rank) else den * extra_den
        for x in row:
            p, q = reduce_pq(x, d)
            out.append([p, q])
    return out


def pq_of_fracs(grid):
    return [[x.numerator, x.denominator] for row in grid for x in row]


def as_fraction(x):
    if isinstance(x, Fraction):
        return x
    if isinstance(x, int):
        return Fraction(x)
    return Fraction(int(x.p), int(x.q))


def np_batch(mats):
    return np.ascontiguousarray(np.array(mats, dtype=np.int32))


__all__ = ["pq_grid_from_num", "pq_of_fracs", "as_fraction", "np_batch", "limbs_to_ints"]
