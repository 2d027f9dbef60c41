"""The multi-modular scheme the CUDA kernels implement (tests/device_model.py mirrors them word for
word) against the exact oracle.  CPU only."""
import random
from fractions import Fraction

from oracle import golden_io, ref_port
from tests import device_model as dm

PRIMES = dm.prime_table(12)
PRECS = [dm.Prime(p) for p in PRIMES]


def modular_rref(A, bar, K):
    m, n = len(A), len(A[0])
    per = [dm.elim_words(A, bar, PRECS[k]) for k in range(K)]
    prof = per[0][2]
    assert all(p[2] == prof for p in per), "bad prime in a test that does not expect one"
    rank = per[0][3]
    N = [[dm.garner_signed([per[k][0][i][j] for k in range(K)], PRIMES[:K]) for j in range(n)] for i in range(m)]
    d = dm.garner_signed([per[k][1] for k in range(K)], PRIMES[:K])
    return N, d, prof, rank


def check_against_oracle(A, bar):
    m, n = len(A), len(A[0])
    amax = max(1, max(abs(x) for r in A for x in r))
    bits = dm.log2_minor_bound(m, bar, n > bar, amax, amax, False, 0)
    K, L = dm.plan_bits(bits)
    N, d, prof, rank = modular_rref(A, bar, K)
    R, piv = ref_port.row_reduce(A, bar)
    assert d != 0
    assert rank == len(piv)
    assert [j for j, s in enumerate(prof) if s != dm.SKIP] == [c for _, c in piv]
    for i in range(m):
        for j in range(n):
            assert Fraction(N[i][j], d) == R[i][j], (A, bar, i, j)
    assert abs(d) < 2 ** (32 * L - 1)
    assert all(abs(x) < 2 ** (32 * L - 1) for r in N for x in r)


def test_mont_roundtrip():
    P = PRECS[0]
    for a in (0, 1, 2, 12345, P.p - 1):
        w = dm.mont_mul(a, P.r2, P)
        assert dm.mont_redc(w, P) == a
    assert dm.mont_mul(P.one, P.one, P) == P.one


def test_prime_table():
    assert PRIMES[0] == 2**31 - 1
    assert all(dm.is_prime_u32(p) and p > 2 ** 30.999 for p in PRIMES)
    assert len(set(PRIMES)) == len(PRIMES)


def test_edge_cases_match_reference_golden():
    g = golden_io.load("edge_small")
    n = 0
    for c in g["rref_cases"][::3]:
        A = c["A"]
        bar = c["bar_col"] or len(A[0]) - 1
        if bar <= 0:
            continue
        check_against_oracle(A, bar)
        n += 1
    assert n > 100


def test_random_shapes():
    rnd = random.Random(7)
    for _ in range(60):
        m, n = rnd.randint(1, 7), rnd.randint(1, 8)
        r = rnd.randint(0, min(m, n))
        B = [[rnd.randint(-5, 5) for _ in range(r)] for _ in range(m)]
        C = [[rnd.randint(-5, 5) for _ in range(n)] for _ in range(r)]
        A = [[sum(B[i][k] * C[k][j] for k in range(r)) for j in range(n)] for i in range(m)]
        check_against_oracle(A, rnd.randint(1, n))


def test_row_norm_bound_is_rigorous_and_never_above_the_plan():
    """The prime count the device takes from the row norms of the matrices at hand (k_row_bound, mirrored by
    dm.row_norm_bits): every exact integer of the elimination stays below the bound, the count never exceeds the
    plan's, and the reconstruction from that many primes is the one from the plan's count.  Shapes with a general
    right block, an identity block ([A|I], incl. singular A), right-hand sides of solves (consistent and not),
    rank-deficient products with a declared maximal rank, zero rows and worst-case rows."""
    rng = random.Random(20260219)
    def ints(m, n, lo):
        return [[rng.randint(-lo, lo) for _ in range(n)] for _ in range(m)]
    def product(m, n, rk, lo):
        B, C = ints(m, rk, lo), ints(rk, n, lo)
        return [[sum(B[i][k] * C[k][j] for k in range(rk)) for j in range(n)] for i in range(m)]
    cases = []
    for _ in range(40):
        m, n = rng.randint(1, 7), rng.randint(1, 8)
        bar = rng.randint(1, n)
        A = ints(m, n, rng.choice([1, 5, 50, 3000]))
        if rng.random() < 0.3:
            A[rng.randrange(m)] = [0] * n
        cases.append(("rref", A, None, bar, 0, False))
    for _ in range(25):
        m = rng.randint(2, 7)
        n = rng.randint(2, 7)
        rk = rng.randint(1, min(m, n))
        A = product(m, n, rk, 5)
        x = [rng.randint(-5, 5) for _ in range(n)]
        b = [sum(A[i][j] * x[j] for j in range(n)) for i in range(m)] if rng.random() < 0.6 else [rng.randint(-9, 9) for _ in range(m)]
        cases.append(("solve", A, b, n, rk, False))
    for _ in range(20):
        m = rng.randint(1, 6)
        A = ints(m, m, rng.choice([1, 5, 100]))
        if rng.random() < 0.3 and m > 1:
            A[-1] = list(A[0])
        cases.append(("inverse", A, None, m, 0, True))
    for _ in range(8):                                     # one large entry sets the declared magnitude, the rest is small
        m = rng.randint(7, 10)
        A = ints(m, m + 1, 5)
        A[rng.randrange(m)][rng.randrange(m)] = 30000
        cases.append(("rref", A, None, m, 0, False))
        B = ints(m, m, 3)
        B[0][0] = 2000
        cases.append(("inverse", B, None, m, 0, True))
    cases.append(("rref", [[7, 7, 7], [7, 7, 7], [7, 7, 7]], None, 3, 0, False))           # every entry at the declared magnitude
    cases.append(("inverse", [[1000, 0, 0], [0, -1000, 0], [0, 0, 1000]], None, 3, 0, True))  # the bound is attained
    fewer = 0
    for op, A, b, bar, max_rank, ident in cases:
        m = len(A)
        full = [list(row) + ([b[i]] if b is not None else []) + ([1 if i == j else 0 for j in range(m)] if ident else [])
                for i, row in enumerate(A)]
        n = len(full[0])
        amax = max(1, max(abs(x) for r in A for x in r))
        bmax = max(1, max(abs(x) for x in b)) if b is not None else (amax if not ident else 1)
        Kp, _ = dm.plan_bits(dm.log2_minor_bound(m, bar, n > bar, amax, bmax, ident, max_rank))
        bits = dm.row_norm_bits(A, b, bar, max_rank, ident)
        Ke = min(Kp, dm.plan_bits(bits)[0])
        Kbig = Kp + 2
        N, d, prof, rank = modular_rref(full, bar, Kbig)
        assert max_rank == 0 or rank <= max_rank
        biggest = max([abs(d)] + [abs(x) for r in N for x in r])
        assert biggest <= 2.0 ** bits * (1 + 1e-9), (op, A, b, bar, biggest, bits)
        N2, d2, prof2, rank2 = modular_rref(full, bar, Ke)
        assert (N2, d2, prof2, rank2) == (N, d, prof, rank), (op, A, b, bar, Ke, Kp)
        fewer += Ke < Kp
    assert fewer >= 5


def test_garner_signed_range():
    ps = PRIMES[:3]
    M = ps[0] * ps[1] * ps[2]
    for x in (0, 1, -1, M // 2, -(M // 2), 123456789123456789, -98765432109876543210):
        assert dm.garner_signed([x % p for p in ps], ps) == x


def test_bad_prime_profile_is_lexicographically_larger():
    # p divides the first pivot candidate: that prime picks a later row (or skips), so its profile
    # compares above the rational one -- the rule k_verify relies on.
    p = PRIMES[0]
    A = [[p, 1], [1, 1]]                    # A[0][0] = 0 mod p, rational profile is [0, 1]
    A = [[a if abs(a) < p else 0 for a in row] for row in A]   # entries are reduced mod p on load
    bad = dm.elim_words(A, 2, PRECS[0])[2]
    good = dm.elim_words([[PRIMES[0], 1], [1, 1]], 2, PRECS[1])[2]
    assert good == [0, 1]
    assert bad > good


def test_inplace_single_prime_inverse_matches_oracle():
    rnd = random.Random(3)
    P = PRECS[0]
    n_sing = 0
    for n in (1, 2, 3, 4, 5, 8):
        for t in range(40):
            A = [[rnd.randint(-5, 5) for _ in range(n)] for _ in range(n)]
            if t % 5 == 0 and n > 1:
                A[rnd.randrange(n)][0] = 0        # exercise row swaps
            if t % 9 == 0 and n > 1:
                A[-1] = list(A[0])                # singular
            got = dm.inverse_inplace_words(A, P)
            want = ref_port.inverse(A)
            if want is None:
                assert got is None
                n_sing += 1
                continue
            adj, det = got
            assert det == ref_port.bareiss_det(A)
            assert [[Fraction(x, det) for x in row] for row in adj] == want
    assert n_sing > 0
    g = golden_io.load("c2_8x8")
    for c in g["cases"][:64] + g["cases"][-8:]:
        got = dm.inverse_inplace_words(c["A"], P)
        want = ref_port.inverse(c["A"])
        assert (got is None) == (want is None)
        if got:
            assert [[Fraction(x, got[1]) for x in row] for row in got[0]] == want


def test_inplace_v2_head_steps_and_folding():
    assert dm.head_steps_for(8, 5) == 3 and dm.head_steps_for(8, 7) == 3 and dm.head_steps_for(8, 8) == 2
    assert dm.head_steps_for(8, 127) == 2 and dm.head_steps_for(8, 128) == 1 and dm.head_steps_for(8, 40000) == 0
    assert dm.head_steps_for(2, 5) == 1 and dm.head_steps_for(1, 5) == 0
    P = PRECS[0]
    x = 123456789
    assert dm.mont_inv_mersenne31(x, P) == dm.mont_pow(x, P.p - 2, P)
    rnd = random.Random(17)
    n_sing = 0
    for P in (PRECS[0], PRECS[1]):
        for n in (1, 2, 3, 4, 5, 6, 7, 8):
            for head in range(0, min(3, n - 1) + 1):
                for t in range(12):
                    A = [[rnd.randint(-5, 5) for _ in range(n)] for _ in range(n)]
                    if t % 4 == 0 and n > 1:
                        A[rnd.randrange(n)][0] = 0
                        A[0][0] = 0
                    if t % 5 == 0 and n > 1:
                        A[-1] = list(A[0])
                    if t % 6 == 1 and n > 2:
                        A[n - 1][n - 1] = 0
                        A[n - 2][n - 2] = 0
                    got = dm.inverse_inplace_v2(A, P, head)
                    want = ref_port.inverse(A)
                    if want is None:
                        assert got is None
                        n_sing += 1
                        continue
                    adj, det = got
                    assert det == ref_port.bareiss_det(A), (n, head, A)
                    assert [[Fraction(x, det) for x in row] for row in adj] == want, (n, head, A)
    assert n_sing > 10


def test_inplace_bareiss_integer_kernel_mirror():
    """tpm_eliminate_bareiss (the default fused 8x8 kernel): exact 32-bit integer Gauss-Jordan with division by the
    previous pivot as a multiplication modulo 2^32; both the 32-bit-step and the all-64-bit instantiation."""
    assert dm.h32_steps_for(8, 5) == 4 and dm.h32_steps_for(8, 9) == 3 and dm.h32_steps_for(4, 5) == 4
    assert dm.h32_steps_for(8, 40000) == 0 and dm.h32_steps_for(1, 5) == 1
    for o in (1, 3, 5, -7, 0x7fffffff, -0x7fffffff, 123456789):
        assert (dm.inv_odd_u32(o) * (o & 0xFFFFFFFF)) & 0xFFFFFFFF == 1
    rnd = random.Random(23)
    n_sing = 0
    for n in (1, 2, 3, 4, 5, 6, 7, 8):
        for h32 in sorted({0, min(n, 4)}):
            for t in range(16):
                A = [[rnd.randint(-5, 5) for _ in range(n)] for _ in range(n)]
                if t % 4 == 0 and n > 1:
                    A[rnd.randrange(n)][0] = 0
                    A[0][0] = 0
                if t % 5 == 0 and n > 1:
                    A[-1] = list(A[0])
                if t % 6 == 1 and n > 2:
                    A[n - 1][n - 1] = 0
                    A[n - 2][n - 2] = 0
                if t % 7 == 2:
                    A = [[2 * x for x in row] for row in A]          # even pivots: shifts in every division
                    A = [[max(-5, min(5, x)) for x in row] for row in A]
                got = dm.inverse_inplace_bareiss(A, h32)
                want = ref_port.inverse(A)
                if want is None:
                    assert got is None
                    n_sing += 1
                    continue
                adj, det = got
                assert det == ref_port.bareiss_det(A), (n, h32, A)
                assert [[Fraction(x, det) for x in row] for row in adj] == want, (n, h32, A)
    assert n_sing > 10
    # entries near the limit of the fused path for n = 4: |a| <= 180 keeps every minor below 2^31
    for t in range(20):
        A = [[rnd.randint(-180, 180) for _ in range(4)] for _ in range(4)]
        got = dm.inverse_inplace_bareiss(A, dm.h32_steps_for(4, 180) if dm.h32_steps_for(4, 180) >= 4 else 0)
        want = ref_port.inverse(A)
        if want is not None:
            assert [[Fraction(x, got[1]) for x in row] for row in got[0]] == want
