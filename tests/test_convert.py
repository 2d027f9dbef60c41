"""Host-side number conversion (linalg_solver_b200/convert.py): limb decoding, CRT basis and rational
reconstruction used by the step trace.  Pure Python, no GPU."""
import random
from fractions import Fraction

import numpy as np

from tests.device_model import prime_table


def _convert():
    # convert.py has no dependency on the CUDA library; load it without importing the package (which needs liblsx.so)
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "linalg_solver_b200", "convert.py")
    spec = importlib.util.spec_from_file_location("lsx_convert", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_limbs_to_ints_signed_multi_limb():
    conv = _convert()
    rnd = random.Random(3)
    for L in (1, 2, 3, 11):
        vals = [0, 1, -1, 2 ** (32 * L - 1) - 1, -(2 ** (32 * L - 1))] + [rnd.randrange(-(2 ** (32 * L - 1)), 2 ** (32 * L - 1)) for _ in range(20)]
        words = np.array([[(v >> (32 * i)) & 0xffffffff for i in range(L)] for v in vals], dtype=np.uint32)
        assert conv.limbs_to_ints(words) == vals


def test_crt_and_rational_reconstruction_round_trip():
    conv = _convert()
    rnd = random.Random(5)
    primes = prime_table(5)
    M, coef = conv.crt_basis(primes)
    bound = 1 << 60                                  # 2 * bound^2 < M (five 31-bit primes: M > 2^154)
    for _ in range(300):
        q = rnd.randrange(1, bound)
        p = rnd.randrange(-bound, bound)
        fr = Fraction(p, q)
        if any(fr.denominator % pr == 0 for pr in primes):
            continue
        x = sum(((fr.numerator * pow(fr.denominator, pr - 2, pr)) % pr) * c for pr, c in zip(primes, coef)) % M
        assert conv.rational_reconstruct(x, M, bound) == fr
    # a value that is not a small fraction is rejected instead of being mis-reconstructed
    assert all(conv.rational_reconstruct(rnd.randrange(M), M, 1 << 20) is None for _ in range(50))
    assert conv.reduce_pq(6, -4) == (-3, 2)
