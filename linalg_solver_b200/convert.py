"""Packed device results -> Python exact numbers.

The C-ABI returns every integer as ``limbs`` little-endian 32-bit words in two's complement
(include/lsx.h).  These helpers are host-side glue, outside any timed region.
"""
from fractions import Fraction
from math import gcd

import numpy as np


def to_numpy(x):
    """numpy view/copy of a result buffer (numpy array, or torch tensor on any device)."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x)


def limbs_to_ints(words):
    """[..., L] words -> nested list of Python ints with the leading shape."""
    w = to_numpy(words)
    L = w.shape[-1]
    lead = w.shape[:-1]
    if L == 1:
        flat = w.view(np.int32).reshape(-1).tolist()
    elif L == 2:
        flat = w.view(np.uint32).reshape(-1, 2).copy().view(np.int64).reshape(-1).tolist()
    else:
        raw = w.view(np.uint32).reshape(-1, L)
        nbytes = 4 * L
        buf = raw.astype("<u4").tobytes()
        flat = [int.from_bytes(buf[i * nbytes:(i + 1) * nbytes], "little", signed=True)
                for i in range(raw.shape[0])]
    return _nest(flat, lead)


def _nest(flat, shape):
    if len(shape) == 0:
        return flat[0]
    if len(shape) == 1:
        return list(flat)
    step = 1
    for s in shape[1:]:
        step *= s
    return [_nest(flat[i * step:(i + 1) * step], shape[1:]) for i in range(shape[0])]


def reduce_pq(num, den):
    """Lowest-terms (p, q) with q > 0 of num/den (den != 0)."""
    if den < 0:
        num, den = -num, -den
    g = gcd(num, den)
    return (num // g, den // g) if g > 1 else (num, den)


def fraction_grid(num_rows, den):
    """Rows of integer numerators over one denominator -> rows of Fractions."""
    return [[Fraction(x, den) for x in row] for row in num_rows]


# ---- residues -> rationals (step traces: every intermediate entry is a quotient of minors) ---------------
def crt_basis(primes):
    """(M, [c_k]) with x = sum r_k * c_k mod M for residues r_k modulo the given primes."""
    M = 1
    for p in primes:
        M *= int(p)
    coef = []
    for p in primes:
        p = int(p)
        Mk = M // p
        coef.append(Mk * pow(Mk % p, p - 2, p))
    return M, coef


def rational_reconstruct(x, M, bound):
    """The fraction a / b with |a|, b <= bound and a = b * x (mod M), or None (Wang's algorithm: the extended
    Euclidean sequence of (M, x) is stopped at the first remainder not above ``bound``).  Unique when 2 bound^2 < M."""
    r0, r1, t0, t1 = M, x % M, 0, 1
    while r1 > bound:
        q = r0 // r1
        r0, r1 = r1, r0 - q * r1
        t0, t1 = t1, t0 - q * t1
    if t1 == 0 or abs(t1) > bound:
        return None
    if t1 < 0:
        r1, t1 = -r1, -t1
    if gcd(r1, t1) != 1:
        return None
    return Fraction(r1, t1)
