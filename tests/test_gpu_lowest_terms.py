"""lsx_lowest_terms (device gcd + exact division) against the host checker convert.reduce_pq: the reference returns
reduced rationals (linalg.py:574, 698-699), the batched kernels return numerators over one common denominator."""
import random

import numpy as np
import pytest

from linalg_solver_b200.convert import limbs_to_ints, reduce_pq
from oracle import golden_io

pytestmark = pytest.mark.gpu


def to_limbs(values, L):
    """Python ints -> [len, L] uint32 two's complement."""
    out = np.zeros((len(values), L), dtype=np.uint32)
    for i, v in enumerate(values):
        out[i] = np.frombuffer(int(v).to_bytes(4 * L, "little", signed=True), dtype="<u4")
    return out


@pytest.fixture(scope="module")
def eng():
    from linalg_solver_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("L", [1, 2, 3, 4, 7, 11, 12, 20, 21, 32])
def test_random_pairs_with_planted_common_factors(eng, L):
    rnd = random.Random(100 + L)
    B, C = 24, 9
    bits = 32 * L - 2
    nums, dens = [], []
    for b in range(B):
        g = rnd.getrandbits(rnd.randint(0, bits // 2)) | 1
        g <<= rnd.randint(0, 7) if b % 3 else 0                  # even common factors too
        d = g * (rnd.getrandbits(max(1, bits - g.bit_length() - 1)) + 1)
        if b % 2:
            d = -d
        d = max(-(1 << bits), min((1 << bits) - 1, d))
        if b == 5:
            d = 0                                                # a flagged matrix: 0 / 0
        dens.append(d)
        row = []
        for i in range(C):
            room = max(1, bits - g.bit_length() - 1)
            x = g * rnd.getrandbits(rnd.randint(1, room)) * rnd.choice([1, -1])
            if i == 0:
                x = 0
            if i == 1:
                x = d                                            # -> 1 / 1
            if i == 2:
                x = -d
            if i == 3:
                x = rnd.getrandbits(bits) - (1 << (bits - 1))    # no planted factor
            row.append(x)
        nums.append(row)
    num = to_limbs([x for row in nums for x in row], L).reshape(B, C, L)
    den = to_limbs(dens, L)
    for container in ("numpy", "torch"):
        if container == "torch":
            import torch
            p, q = eng.lowest_terms(torch.from_numpy(num.view(np.int32)).cuda(), torch.from_numpy(den.view(np.int32)).cuda())
        else:
            p, q = eng.lowest_terms(num, den)
        P, Q = limbs_to_ints(p), limbs_to_ints(q)
        for b in range(B):
            for i in range(C):
                want = (0, 0) if dens[b] == 0 else reduce_pq(nums[b][i], dens[b])
                assert (P[b][i], Q[b][i]) == want, (L, b, i, nums[b][i], dens[b])


def test_golden_inverses_reduced_on_the_device(eng):
    """Config 2 and config 4 goldens: the reference's inverse entries (p, q) straight from the device."""
    g = golden_io.load("c2_8x8")
    cases = [c for c in g["cases"] if c.get("inverse")][:64]
    A = np.array([c["A"] for c in cases], dtype=np.int32)
    res = eng.inverse_batch(A, a_abs_max=5)
    p, q = eng.lowest_terms(res.adj, res.det)
    P, Q = limbs_to_ints(p), limbs_to_ints(q)
    for k, c in enumerate(cases):
        assert [[P[k][i][j], Q[k][i][j]] for i in range(8) for j in range(8)] == c["inverse"]
    g4 = golden_io.load("c4_64x64")
    inv_cases = [c for c in g4["inverse_cases"]] if "inverse_cases" in g4 else [c for c in g4["cases"] if "inverse_sha" in c]
    A4 = np.array([c["A"] for c in inv_cases[:8]], dtype=np.int32)
    r4 = eng.inverse_batch(A4, a_abs_max=5)
    p4, q4 = eng.lowest_terms(r4.adj, r4.det)
    P4, Q4 = limbs_to_ints(p4), limbs_to_ints(q4)
    for k, c in enumerate(inv_cases[:8]):
        flat = [(P4[k][i][j], Q4[k][i][j]) for i in range(64) for j in range(64)]
        assert golden_io.digest_pq(flat) == c["inverse_sha"]


def test_full_size_config4_lowest_terms_properties(eng):
    """2^14 of config 4's 64x64 inverses (6.7e7 eleven-limb fractions): the reduction is idempotent (its output is
    already in lowest terms, so reducing p / q again changes nothing: gcd(p, q) = 1), q > 0, p / q = num / den
    (cross products modulo a prime outside the table), and a sample equals the host checker."""
    import torch
    B = 1 << 14
    rng = np.random.Generator(np.random.PCG64(77))
    A = torch.from_numpy(rng.integers(-5, 6, size=(B, 64, 64), dtype=np.int32)).cuda()
    res = eng.inverse_batch(A, a_abs_max=5)
    p, q = eng.lowest_terms(res.adj, res.det)
    L = p.shape[-1]
    # idempotence, per matrix entry i: lowest_terms(p[i], q[i]) == (p[i], q[i]); den varies per entry, so regroup
    pe, qe = p.reshape(B * 4096, 1, L), q.reshape(B * 4096, L)
    p2, q2 = eng.lowest_terms(pe, qe)
    assert torch.equal(p2.reshape(p.shape), p) and torch.equal(q2.reshape(q.shape), q)
    assert bool((q[..., L - 1] >= 0).all())                       # sign word of q: positive (or the 0 / 0 of a singular matrix)
    # cross products modulo a 31-bit prime that is not in the table: p * den == num * q
    P = 1000000007                                                # a prime far below the table (the 2048 largest primes under 2^31)
    def mod_p(x):                                                 # [..., L] int32 words (two's complement) -> residues (int64)
        w = x.to(torch.int64) & 0xffffffff
        acc = torch.zeros(x.shape[:-1], dtype=torch.int64, device=x.device)
        for l in range(L - 1, -1, -1):
            acc = (acc * ((1 << 32) % P) + w[..., l]) % P
        neg = x[..., L - 1] < 0
        return torch.where(neg, (acc - pow(2, 32 * L, P)) % P, acc)
    lhs = (mod_p(p) * mod_p(res.det)[:, None, None]) % P
    rhs = (mod_p(res.adj) * mod_p(q)) % P
    assert torch.equal(lhs, rhs)
    adj, det = limbs_to_ints(res.adj[:2]), limbs_to_ints(res.det[:2])
    Ps, Qs = limbs_to_ints(p[:2]), limbs_to_ints(q[:2])
    for k in range(2):
        for i in range(64):
            for j in range(64):
                assert (Ps[k][i][j], Qs[k][i][j]) == reduce_pq(adj[k][i][j], det[k])
