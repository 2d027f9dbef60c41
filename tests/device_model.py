"""Host model of the DEVICE algorithm (word-for-word what the CUDA kernels compute).

Test infrastructure: lets the CPU suite check the multi-modular scheme (division-free
uniform-scale Gauss-Jordan on Montgomery words, one inversion per matrix and prime,
Garner CRT to signed integers, pivot-profile agreement) against the exact oracle
without a GPU.  The CUDA code in linalg_solver_b200/csrc mirrors these functions;
names match (mont_redc, elim_words, garner_signed, plan_bits).
"""
import math

R = 1 << 32
MASK = R - 1
SKIP = 31


def is_prime_u32(n):
    if n < 2:
        return False
    for q in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % q == 0:
            return n == q
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7):          # deterministic below 3,215,031,751
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def prime_table(count, start=(1 << 31) - 1):
    """Primes below 2^31 in descending order (same table as lsx_primes.cpp)."""
    out = []
    n = start
    while len(out) < count:
        if is_prime_u32(n):
            out.append(n)
        n -= 2
    return out


class Prime:
    def __init__(self, p):
        self.p = p
        self.pinv = (-pow(p, -1, R)) % R        # -p^{-1} mod 2^32
        self.one = R % p                        # word of value 1
        self.r2 = (R * R) % p                   # word of value R


def mont_redc(t, P):
    m = ((t & MASK) * P.pinv) & MASK
    u = (t + m * P.p) >> 32
    return u - P.p if u >= P.p else u


def mont_mul(a, b, P):
    return mont_redc(a * b, P)


def mont_pow(a, e, P):
    acc = P.one
    for bit in bin(e)[2:]:
        acc = mont_mul(acc, acc, P)
        if bit == "1":
            acc = mont_mul(acc, a, P)
    return acc


def elim_words(A, bar, P, left_done_skip=False):
    """One matrix, one prime.  A: m x n Python ints with |a| < p.

    Returns (N, d, profile, rank): N[i][j] = d * RREF[i][j] mod p as plain residues,
    d = sign * prod(pivots) mod p (determinant of the pivot minor), profile[j] = source
    row chosen for column j (SKIP when the column has no pivot).
    Words are loaded RAW (w = a mod p, i.e. value a/R): every value carries the same
    constant factor 1/R, which the factors Gw (pivot rows) and G2w (non-pivot rows)
    take out again at the end.
    """
    m, n = len(A), len(A[0])
    p = P.p
    W = [[a % p for a in row] for row in A]
    S, Q, X = P.one, P.one, 1
    pi, neg = 0, False
    profile = []
    for j in range(bar):
        src = next((r for r in range(pi, m) if W[r][j] != 0), None)
        if src is None:
            profile.append(SKIP)
            continue
        profile.append(src)
        if src != pi:
            W[pi], W[src] = W[src], W[pi]
            neg = not neg
        piv = W[pi][j]
        prow = list(W[pi])
        for r in range(m):
            x = S if r == pi else piv
            y = 0 if r == pi else p - W[r][j]
            c0 = j + 1 if left_done_skip else 0
            for c in range(c0, n):
                W[r][c] = mont_redc(x * W[r][c] + y * prow[c], P)
        Q = mont_mul(Q, S, P)
        S = mont_mul(S, piv, P)
        X = mont_mul(X, P.r2, P)
        pi += 1
    qinv = mont_pow(Q, p - 2, P)
    Gw = mont_mul(qinv, X, P)
    if neg and Gw:
        Gw = p - Gw
    G2w = mont_mul(Gw, P.r2, P)
    N = [[mont_mul(Gw if r < pi else G2w, W[r][c], P) for c in range(n)] for r in range(m)]
    d = mont_mul(Gw, S, P)
    return N, d, profile, pi


def garner_signed(res, primes):
    """Residues (one per prime) -> the unique integer in (-M/2, M/2)."""
    K = len(primes)
    v = []
    for j in range(K):
        pj = primes[j]
        t = res[j] % pj
        for i in range(j):
            t = (t - v[i]) * pow(primes[i], -1, pj) % pj
        v.append(t)
    neg = False
    for i in reversed(range(K)):
        h = (primes[i] - 1) // 2
        if v[i] != h:
            neg = v[i] > h
            break
    x = 0
    for i in reversed(range(K)):
        x = x * primes[i] + v[i]
    if neg:
        M = 1
        for q in primes:
            M *= q
        x -= M
    return x


def log2_minor_bound(m, bar, has_right, a_abs, b_abs, right_identity, max_rank):
    """log2 of the largest |minor| the outputs can be (Hadamard), see DESIGN.md.

    Every output integer (common denominator d, numerators d*R[i][j]) is a minor of
    [A|B] of size s <= smax that uses at most one right-block column.
    """
    a = max(1, int(a_abs))
    b = max(1, int(b_abs))
    r = min(m, bar)
    if max_rank and max_rank > 0:
        r = min(r, max_rank)
    best = r * (0.5 * math.log2(r) + math.log2(a)) if r > 0 else 0.0
    if has_right:
        s = min(m, r + 1)
        if right_identity:
            t = s - 1
            cand = t * (0.5 * math.log2(t) + math.log2(a)) if t > 0 else 0.0
        else:
            cand = (s - 1) * (0.5 * math.log2(s) + math.log2(a)) + 0.5 * math.log2(s) + math.log2(b)
        best = max(best, cand)
    return best


PRIME_BITS = 30.999        # every table prime is > 2^30.999


def plan_bits(log2_bound):
    need = log2_bound + 1.0 + 1e-6          # sign bit + rounding slack
    K = max(1, math.ceil(need / PRIME_BITS))
    L = max(1, math.ceil(need / 32.0))
    return K, L


def row_norm_bits(A, b, bar, max_rank=0, right_identity=False):
    """Mirror of k_row_bound / bound_word (lsx_tile.cu): log2 of the Hadamard bound over the row norms of the matrix at
    hand for every integer an elimination of [A | right part] returns, in the kernel's 1/256-bit fixed point.  A holds
    the n_in stored columns (a general right block included), b is the right-hand side of a solve or None, and
    right_identity says that an identity block follows A.  plan_bits() of the result is the prime count the device runs
    (never above the plan's)."""
    m, n_in = len(A), len(A[0])
    n = n_in + (1 if b is not None else 0) + (m if right_identity else 0)
    lg = []
    for row in A:
        s = sum(x * x for x in row) + (1 if right_identity else 0)
        lg.append(0.5 * math.log2(s) * (1.0 + 1e-12) + 1e-4 if s > 1 else 0.0)
    r = min(m, bar)
    if 0 < max_rank < r:
        r = max_rank
    r_top = r + 1 if (n > bar and r < m) else r
    top = sorted(lg, reverse=True)
    if b is None:
        tot = sum(top[:r_top])
    else:
        b1 = sum(abs(x) for x in b)
        lb = math.log2(b1) * (1.0 + 1e-12) if b1 > 1 else 0.0
        tot = max(sum(top[:r]), lb + sum(top[:r_top - 1]))
    return (math.ceil(tot * 256.0) + 1) / 256.0


def inverse_inplace_words(A, P):
    """Mirror of the fused small-matrix kernel (lsx_small.cu, k_inv_tpm): in-place uniform-scale
    Gauss-Jordan inversion of one n x n matrix modulo ONE prime larger than every minor of A.

    Returns (adj, det) as exact integers, or None when A is singular.  The column produced at step j
    is stored in the pivot column's slot; `unit[r]` tracks which column of the (virtual) right block
    still holds row r's unit entry, which gives the column permutation undone at the end.
    """
    n = len(A)
    p = P.p
    W = [[a % p for a in row] for row in A]
    S, Q, X, D = P.one, P.one, 1, 1          # D: word of the right block's diagonal for unpivoted rows
    neg = False
    unit = list(range(n))
    outcol = [0] * n
    for j in range(n):
        src = next((r for r in range(j, n) if W[r][j] != 0), None)
        if src is None:
            return None
        if src != j:
            W[j], W[src] = W[src], W[j]
            unit[j], unit[src] = unit[src], unit[j]
            neg = not neg
        outcol[j] = unit[j]
        piv = W[j][j]
        prow = list(W[j])
        for r in range(n):
            if r == j:
                for c in range(n):
                    W[r][c] = mont_mul(S, D, P) if c == j else mont_mul(S, prow[c], P)
            else:
                f = W[r][j]
                y = p - f if f else 0
                for c in range(n):
                    W[r][c] = mont_mul(y, D, P) if c == j else mont_redc(piv * W[r][c] + y * prow[c], P)
        Q = mont_mul(Q, S, P)
        S = mont_mul(S, piv, P)
        D = mont_mul(D, piv, P)
        X = mont_mul(X, P.r2, P)
    qinv = mont_pow(Q, p - 2, P)
    Gw = mont_mul(qinv, X, P)
    if neg and Gw:
        Gw = p - Gw
    half = p >> 1
    adj = [[0] * n for _ in range(n)]
    for r in range(n):
        for j in range(n):
            v = mont_mul(Gw, W[r][j], P)
            adj[r][outcol[j]] = v - p if v > half else v
    det = sum(A[0][c] * adj[c][0] for c in range(n))
    return adj, det


def head_steps_for(n, a_abs_max):
    """Number of leading pivot steps that can run on plain int32 (no reduction): entries grow like
    B -> 2 B^2 per step (pivot rows are left unscaled, see inverse_inplace_v2)."""
    h, B = 0, max(1, int(a_abs_max))
    while h < min(3, n - 1):
        B = 2 * B * B
        if B >= 2 ** 31:
            break
        h += 1
    return h


def mont_inv_mersenne31(a, P):
    """a^(p-2) for p = 2^31 - 1 by an addition chain: 30 squarings + 8 multiplications."""
    def sqn(x, k):
        for _ in range(k):
            x = mont_mul(x, x, P)
        return x
    x2 = mont_mul(sqn(a, 1), a, P)            # a^(2^2 - 1)
    x4 = mont_mul(sqn(x2, 2), x2, P)          # 2^4 - 1
    x8 = mont_mul(sqn(x4, 4), x4, P)
    x16 = mont_mul(sqn(x8, 8), x8, P)
    x24 = mont_mul(sqn(x16, 8), x8, P)
    x28 = mont_mul(sqn(x24, 4), x4, P)
    x29 = mont_mul(sqn(x28, 1), a, P)         # 2^29 - 1
    return mont_mul(sqn(x29, 2), a, P)        # 4 (2^29 - 1) + 1 = 2^31 - 3


def inverse_inplace_v2(A, P, head):
    """Mirror of k_inv_tpm (lsx_small.cu), second version.

    In-place Gauss-Jordan inversion of one n x n matrix modulo one prime with
      * `head` leading pivot steps on plain integers (int32 range guaranteed by head_steps_for),
      * pivot rows left UNSCALED (row k then lacks the factor sigma_k = prod_{i<k} piv_i, which is
        put back by the per-row multiplier at the end),
      * the final scaling folded into the multipliers of the last pivot step, so the single modular
        inversion happens before that step.
    Returns (adj, det) or None (singular).
    """
    n = len(A)
    p = P.p
    W = [[int(a) for a in row] for row in A]
    unit = list(range(n))
    outcol = [0] * n
    neg = False
    sig = 1                       # sigma_j = product of the pivots so far (head: exact integer)
    cw = [0] * n                  # per-row multiplier word (deficit of pivot row k)
    qh = 1                        # head: product of sigma_k, k < head (exact integer, reduced mod p)

    def pivot(j, zero):
        nonlocal neg
        src = next((r for r in range(j, n) if W[r][j] != zero), None)
        if src is None:
            return False
        if src != j:
            W[j], W[src] = W[src], W[j]
            unit[j], unit[src] = unit[src], unit[j]
            neg = not neg
        outcol[j] = unit[j]
        return True

    for j in range(head):
        if not pivot(j, 0):
            return None
        piv = W[j][j]
        prow = list(W[j])
        for r in range(n):
            if r == j:
                continue
            f = W[r][j]
            for c in range(n):
                W[r][c] = -f * sig if c == j else piv * W[r][c] - f * prow[c]
                assert abs(W[r][c]) < 2 ** 31
        W[j][j] = sig
        cw[j] = mont_mul(sig % p, P.r2, P)
        qh = qh * sig % p
        sig = sig * piv
        assert abs(sig) < 2 ** 31
    # ---- switch to Montgomery words: raw load, S = D = sigma_h raw ----
    W = [[x % p for x in row] for row in W]
    S = sig % p
    Q = mont_mul(qh, P.r2, P)
    for j in range(head, n):
        last = j == n - 1
        if not pivot(j, 0):
            return None
        piv = W[j][j]
        prow = list(W[j])
        cw[j] = S
        Q = mont_mul(Q, S, P)
        if last:
            qinv = mont_inv_mersenne31(Q, P) if p == 2 ** 31 - 1 else mont_pow(Q, p - 2, P)
            if neg:
                qinv = p - qinv
        for r in range(n):
            if r == j:
                if last:
                    g = mont_mul(qinv, cw[r], P)
                    for c in range(n):
                        W[r][c] = mont_mul(g, S if c == j else prow[c], P)
                else:
                    W[r][j] = S
                continue
            f = W[r][j]
            y = p - f if f else 0
            x = piv
            if last:
                g = mont_mul(qinv, cw[r], P)
                x, y = mont_mul(g, x, P), mont_mul(g, y, P)
            for c in range(n):
                W[r][c] = mont_mul(y, S, P) if c == j else mont_redc(x * W[r][c] + y * prow[c], P)
        S = mont_mul(S, piv, P)
    half = p >> 1
    adj = [[0] * n for _ in range(n)]
    for r in range(n):
        for j in range(n):
            v = W[r][j]
            adj[r][outcol[j]] = v - p if v > half else v
    det = sum(A[0][c] * adj[c][0] for c in range(n))
    return adj, det


# ---- mirror of tpm_eliminate_bareiss (lsx_inv_small.cuh): the fused small inverse over the integers ----------------
_M32 = 0xFFFFFFFF


def _s32(x):
    x &= _M32
    return x - (1 << 32) if x & 0x80000000 else x


def inv_odd_u32(o):
    """o^-1 modulo 2^32 for odd o: (3 o) xor 2 is right to 5 bits, three Newton steps."""
    o &= _M32
    x = ((o * 3) ^ 2) & _M32
    for _ in range(3):
        x = (x * ((2 - o * x) & _M32)) & _M32
    return x


def h32_steps_for(n, a_abs_max):
    """Leading pivot steps whose t = piv * w - f * prow fits 32 bits: 2 M^2 < 2^31 for the Hadamard bound M of the
    minors of order j + 1 (lsx_small.cu::h32_steps_for)."""
    import math
    h = 0
    for j in range(n):
        k = j + 1
        bits = log2_minor_bound(k, k, False, a_abs_max, 0, False, k)
        if 1.0 + 2.0 * bits >= 31.0 - 1e-6:
            break
        h = j + 1
    return h


def inverse_inplace_bareiss(A, h32):
    """One-step fraction-free (Bareiss) Gauss-Jordan in place, on 32-bit two's-complement words exactly as the kernel
    does it: the exact division by the previous pivot d = 2^s * o is ((t >> s) * o^-1) mod 2^32, with t formed in 32
    bits for the first h32 steps and in 64 bits afterwards.  Returns (adj, det) or None (singular)."""
    n = len(A)
    W = [[int(a) & _M32 for a in row] for row in A]
    unit = list(range(n))
    outcol = [0] * n
    neg = False
    dprev, dinv, dsh = 1, 1, 0
    piv = 0
    for j in range(n):
        last = j == n - 1
        src = next((r for r in range(j, n) if W[r][j] != 0), None)
        if src is None:
            return None
        if src != j:
            W[j], W[src] = W[src], W[j]
            unit[j], unit[src] = unit[src], unit[j]
            neg = not neg
        outcol[j] = unit[j]
        piv = W[j][j]
        prow = list(W[j])
        flip = last and neg
        minv = (-dinv) & _M32 if flip else dinv
        for r in range(n):
            if r == j:
                continue
            f = W[r][j]
            nf = (-f) & _M32
            for c in range(n):
                if c == j:
                    W[r][c] = f if flip else nf
                elif j == 0:
                    t = (piv * W[r][c] + nf * prow[c]) & _M32
                    W[r][c] = (-t) & _M32 if flip else t
                elif j < h32:
                    t = (piv * W[r][c] + nf * prow[c]) & _M32
                    exact = _s32(piv) * _s32(W[r][c]) - _s32(f) * _s32(prow[c])
                    assert _s32(t) == exact, "32-bit step overflowed"
                    W[r][c] = ((_s32(t) >> dsh) * minv) & _M32
                else:
                    t = _s32(piv) * _s32(W[r][c]) + _s32(nf) * _s32(prow[c])
                    assert -(1 << 63) <= t < (1 << 63)
                    assert t % _s32(dprev) == 0, "Bareiss division must be exact"
                    W[r][c] = (((t >> dsh) & _M32) * minv) & _M32
        W[j][j] = dprev
        if flip:
            W[j] = [(-x) & _M32 for x in W[j]]
        if not last:
            dprev = piv
            sp = _s32(piv)
            dsh = ((sp & -sp).bit_length() - 1) & 31
            dinv = inv_odd_u32(sp >> dsh)
    det = _s32((-piv) & _M32 if neg else piv)
    adj = [[0] * n for _ in range(n)]
    for r in range(n):
        for j in range(n):
            adj[r][outcol[j]] = _s32(W[r][j])
    return adj, det
