"""CPU oracle: exact-rational restatement of the reference's elimination path.

TEST INFRASTRUCTURE ONLY -- never imported by ``linalg_solver_b200`` (the product
path fails loudly when the CUDA library is missing).  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs.

Parity status: PINNED against outputs of the reference itself.  The reference
has no tests or golden vectors for this path, so
``oracle/gen_golden.py`` runs the unmodified reference Python (imported from
/root/reference with the ``oracle/standin`` module for its Rust dependency) on
seeded ``RandomMatrixBuilder`` inputs and stores the results under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below
against those files.

Arithmetic is ``fractions.Fraction`` (exact); each function cites the reference
lines it restates.  The reference's LaTeX frames (linalg.py:544,553,577,598,623)
are presentation only and are not produced here.
"""
from fractions import Fraction
from math import gcd


def _frac_rows(items):
    return [[Fraction(x) for x in row] for row in items]


def row_reduce(items, bar_col=None):
    """Gauss-Jordan of reference linalg.py:534-630.

    Returns ``(R, pivots)``: R is the reduced m x n grid of Fractions, pivots the
    list of (row, col).  Rule set restated from the reference:
      * ``bar_col`` falsy (None or 0) means n-1                   (linalg.py:543)
      * pivot = entry at (pi, pj) if non-zero, else the FIRST lower row with a
        non-zero entry is swapped up, else the column is skipped  (548-567)
      * pivot row is divided by the pivot from column pj on       (569-575)
      * rows below are eliminated                                 (587-596)
      * after the forward sweep, pivots are visited in reverse and rows above
        are eliminated                                            (611-621)
    Only columns < bar_col are searched for pivots; columns >= bar_col just
    receive the row operations.
    """
    A = _frac_rows(items)
    m, n = len(A), len(A[0])
    bar = bar_col or n - 1
    pivots = []
    pi = pj = 0
    while pi < m and pj < bar:
        if A[pi][pj] == 0:
            src = next((i for i in range(pi + 1, m) if A[i][pj] != 0), None)
            if src is None:
                pj += 1
                continue
            A[pi], A[src] = A[src], A[pi]
        lead = A[pi][pj]
        if lead != 1:
            prow = A[pi]
            for j in range(pj, n):
                prow[j] = prow[j] / lead
        prow = A[pi]
        for k in range(pi + 1, m):
            f = A[k][pj]
            if f != 0:
                rk = A[k]
                for j in range(pj, n):
                    rk[j] -= f * prow[j]
        pivots.append((pi, pj))
        pi += 1
        pj += 1
    for r, c in reversed(pivots):
        prow = A[r]
        for k in range(r):
            f = A[k][c]
            if f != 0:
                rk = A[k]
                for j in range(c, n):
                    rk[j] -= f * prow[j]
    return A, pivots


# step descriptions of the reference (linalg.py:556-559, 581, 603, 627): part of its observable output
STEP_SWAP = r"Výměna řádků $R_{%d}$ a $R_{%d}$"
STEP_NORM = r"Normalizace pivotního řádku %s"
STEP_ELIM_DOWN = r"Eliminace prvků pod pivotem ve sloupci %s"
STEP_ELIM_UP = r"Eliminace nad pivotem ve sloupci %s"


def row_reduce_trace(items, bar_col=None):
    """``row_reduce`` with the reference's step log (linalg.py:544-629).

    Returns ``(R, pivots, frames, steps)``: ``frames[0]`` is the input, ``frames[t + 1]`` the matrix after
    ``steps[t] = (label, description)``.  A step is recorded exactly when the reference records one: a row swap
    (S, 551-562), a normalisation that changed the pivot row, i.e. pivot != 1 (N, 569-584), an elimination below
    with at least one non-zero factor (E, 586-606) and, in reverse pivot order, an elimination above with at
    least one non-zero factor (E, 611-628).  Labels are the kind letter followed by the running step number.
    """
    A = _frac_rows(items)
    m, n = len(A), len(A[0])
    bar = bar_col or n - 1
    frames = [[row[:] for row in A]]
    steps = []

    def record(kind, text):
        steps.append(("%s%d" % (kind, len(steps)), text))
        frames.append([row[:] for row in A])

    pivots = []
    pi = pj = 0
    while pi < m and pj < bar:
        if A[pi][pj] == 0:
            src = next((i for i in range(pi + 1, m) if A[i][pj] != 0), None)
            if src is None:
                pj += 1
                continue
            A[pi], A[src] = A[src], A[pi]
            record("S", STEP_SWAP % (pi + 1, src + 1))
        lead = A[pi][pj]
        if lead != 1:
            for j in range(pj, n):
                A[pi][j] = A[pi][j] / lead
            record("N", STEP_NORM % (pi + 1))
        hit = False
        for k in range(pi + 1, m):
            f = A[k][pj]
            if f != 0:
                hit = True
                for j in range(pj, n):
                    A[k][j] -= f * A[pi][j]
        if hit:
            record("E", STEP_ELIM_DOWN % (pj + 1))
        pivots.append((pi, pj))
        pi += 1
        pj += 1
    for r, c in reversed(pivots):
        hit = False
        for k in range(r):
            f = A[k][c]
            if f != 0:
                hit = True
                for j in range(c, n):
                    A[k][j] -= f * A[r][j]
        if hit:
            record("E", STEP_ELIM_UP % (c + 1))
    return A, pivots, frames, steps


def forward_profile(items, bar_col):
    """Forward sweep only; returns (pivots, source_rows, sign, pivot_values).

    ``source_rows[k]`` is the row index that was swapped into position k when
    pivot k was chosen (== k when no swap); ``sign`` is (-1)^(#swaps);
    ``pivot_values`` are the lead entries before normalisation, so that
    sign * prod(pivot_values) is the determinant of the pivot minor
    (the quantity the device path uses as common denominator).
    Same loop as linalg.py:547-609.
    """
    A = _frac_rows(items)
    m, n = len(A), len(A[0])
    pivots, srcs, vals = [], [], []
    sign = 1
    pi = pj = 0
    while pi < m and pj < bar_col:
        if A[pi][pj] == 0:
            src = next((i for i in range(pi + 1, m) if A[i][pj] != 0), None)
            if src is None:
                pj += 1
                continue
            A[pi], A[src] = A[src], A[pi]
            sign = -sign
            srcs.append(src)
        else:
            srcs.append(pi)
        lead = A[pi][pj]
        vals.append(lead)
        prow = A[pi]
        for j in range(pj, n):
            prow[j] = prow[j] / lead
        for k in range(pi + 1, m):
            f = A[k][pj]
            if f != 0:
                rk = A[k]
                for j in range(pj, n):
                    rk[j] -= f * prow[j]
        pivots.append((pi, pj))
        pi += 1
        pj += 1
    return pivots, srcs, sign, vals


def rank(items):
    """Rank = number of pivots when every column may hold a pivot.

    The reference delegates to ``sympy.Matrix.rank`` (linalg.py:745-747); the
    value is the pivot count of the same elimination with bar_col = n.
    """
    n = len(items[0])
    if n == 0:
        return 0
    return len(forward_profile(items, n)[0])


def determinant(items):
    """det = sign * product of forward-sweep pivots, 0 when rank < n.

    The reference's own routes (Rust-planned executor linalg.py:204-207, legacy
    n! sum 264-345) are infeasible for dense n >= 9; a determinant is unique, so
    the oracle takes it from the elimination of linalg.py:547-609.  n = 0 -> 1
    and n = 1 -> the entry as in linalg.py:197-201.
    """
    n = len(items)
    if n == 0:
        return Fraction(1)
    if len(items[0]) != n:
        raise ValueError("Determinant requires a square matrix")
    if n == 1:
        return Fraction(items[0][0])
    pivots, _, sign, vals = forward_profile(items, n)
    if len(pivots) < n:
        return Fraction(0)
    d = Fraction(sign)
    for v in vals:
        d *= v
    return d


def bareiss_det(items):
    """Fraction-free (Bareiss) integer determinant for larger n.

    Independent of the Fraction sweep above; used as a cross-check and as the
    exact oracle at n = 64..512 where Fractions are too slow.
    """
    M = [[int(x) for x in row] for row in items]
    n = len(M)
    if n == 0:
        return 1
    sign, prev = 1, 1
    for k in range(n - 1):
        if M[k][k] == 0:
            src = next((i for i in range(k + 1, n) if M[i][k] != 0), None)
            if src is None:
                return 0
            M[k], M[src] = M[src], M[k]
            sign = -sign
        pk = M[k][k]
        rowk = M[k]
        for i in range(k + 1, n):
            rowi = M[i]
            f = rowi[k]
            for j in range(k + 1, n):
                rowi[j] = (pk * rowi[j] - f * rowk[j]) // prev
            rowi[k] = 0
        prev = pk
    return sign * M[n - 1][n - 1]


def inverse(items):
    """Inverse via [A|I] with bar_col = n (reference linalg.py:704-743).

    Returns the n x n grid of Fractions, or None where the reference returns
    ``Matrix.NoSolution()`` (left block is not the identity, 725-737).  The
    default sympy route (696-701) yields the same values for a regular matrix.
    """
    n = len(items)
    if any(len(r) != n for r in items):
        raise ValueError("Matrix must be square to invert.")
    aug = [list(items[i]) + [1 if i == j else 0 for j in range(n)] for i in range(n)]
    R, _ = row_reduce(aug, bar_col=n)
    for i in range(n):
        for j in range(n):
            if R[i][j] != (1 if i == j else 0):
                return None
    return [row[n:] for row in R]


def is_inconsistent(R, nvars, bar_col):
    """Reference linalg.py:913-934: a zero left row with non-zero rhs."""
    return any(
        all(row[j] == 0 for j in range(nvars)) and row[bar_col] != 0 for row in R
    )


def affine_from_rref(R, pivots, nvars, bar_col):
    """Reference linalg.py:937-999 (logged route).

    particular[pivot col] = rhs of that pivot row, free variables 0 (960-966);
    one generator per free column f in ascending order with gen[f] = 1 and
    gen[pivot col of row i] = -R[i][f] (973-983).  Returns (particular,
    generator_list) where generator_list is a list of length-n vectors
    (possibly empty; the reference then hands back ``None`` for the matrix).
    """
    piv_col_of_row = {r: c for r, c in pivots}
    pivot_cols = {c for _, c in pivots}
    free = [j for j in range(nvars) if j not in pivot_cols]
    particular = [Fraction(0)] * nvars
    for r, c in piv_col_of_row.items():
        particular[c] = R[r][bar_col]
    gens = []
    for f in free:
        g = [Fraction(0)] * nvars
        g[f] = Fraction(1)
        for r, c in piv_col_of_row.items():
            g[c] = -R[r][f]
        gens.append(g)
    return particular, gens


def find_preimage_of(items, vec):
    """Solve A x = vec.

    Logged route of reference linalg.py:648-680: row_reduce([A|vec], bar = n),
    inconsistency test, extraction.  Returns None for NoSolution, else
    (particular, generators) with generators in ASCENDING free-column order
    (use ``sympy_generator_order`` for the default route's column order).
    """
    m = len(items)
    if m != len(vec):
        raise ValueError("Matrix dimensions must match")
    n = len(items[0])
    aug = [list(items[i]) + [vec[i]] for i in range(m)]
    R, pivots = row_reduce(aug, bar_col=n)
    if is_inconsistent(R, n, n):
        return None
    return affine_from_rref(R, pivots, n, n)


def sympy_generator_order(k):
    """Column order of the default (sympy.linsolve) route, linalg.py:895.

    The reference sorts linsolve's free symbols tau0..tau{k-1} by ``str``; tau_i
    belongs to the i-th free column, so for k > 10 the generator columns come
    out as 0, 1, 10, 11, ..., 2, ...  Returns the permutation as a list of
    ascending-order positions.
    """
    return sorted(range(k), key=lambda i: "tau%d" % i)


def kernel(items):
    """Reference linalg.py:749-756: find_preimage_of(zero vector)."""
    return find_preimage_of(items, [0] * len(items))


def common_denominator_form(R, pivots, srcs_sign_vals):
    """Helper for tests: d * R as integers, d = sign * prod(pivot values)."""
    _, sign, vals = srcs_sign_vals
    d = Fraction(sign)
    for v in vals:
        d *= v
    out = []
    for row in R:
        out.append([x * d for x in row])
    return d, out


def as_pq(x):
    """Canonical (numerator, denominator) pair, denominator > 0, lowest terms."""
    if isinstance(x, Fraction):
        return x.numerator, x.denominator
    if isinstance(x, int):
        return x, 1
    p, q = int(x.p), int(x.q)  # sympy Rational / Integer
    g = gcd(p, q)
    if q < 0:
        p, q = -p, -q
    return p // g, q // g
