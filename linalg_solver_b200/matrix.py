"""``Matrix`` / ``Matrix.AffineSubspace`` / ``Matrix.NoSolution`` with the reference's method
surface for the elimination path, served by the GPU engine.

Mirrors reference linalg_solver/linalg.py: container and validation (11-58), ``zero/identity/
diagonal/new_vector/transpose`` (409-422, 482-489), ``AffineSubspace`` (491-522), ``NoSolution``
(524-532), ``row_reduce`` (534-630), ``find_preimage_of`` (632-680, 870-999), ``inverse`` (682-743),
``rank`` (745-747), ``kernel`` (749-756), the entry of ``determinant`` (183-207) and the eigen-callers that
sit on them (``eigenvalues`` 424-480, ``find_eigenspace`` 758-770, ``diagonalize`` 833-863).  Same names,
argument meaning, return shapes and error texts; the arithmetic runs in liblsx (there is no CPU
fallback) and the reference's LaTeX ``Logger`` output is bypassed: ``row_reduce`` returns empty
lists for its two log results, which the reference's own callers accept (linalg.py:1009-1010).

Entries may be Python ints, ``fractions.Fraction`` or ``sympy`` rationals.  The device works on
machine integers, so a matrix with fractional entries is multiplied by the common denominator
first and the results are scaled back (exactly).  Plain ``int`` input is treated as exact
rationals: the reference itself degrades to floats there (true division, linalg.py:574), which is
why its own driver rationalises inputs first (main.py:20-31).
"""
from fractions import Fraction
from math import gcd
from typing import Any, Callable, Iterator, List, Tuple

import numpy as np

from . import _lib
from .convert import limbs_to_ints, reduce_pq
from .engine import default_engine

_INT32_MAX = 2**31 - 1


def _pq_of(x):
    """(numerator, denominator) of an exact entry, or raise TypeError."""
    if isinstance(x, bool):
        return int(x), 1
    if isinstance(x, int):
        return x, 1
    if isinstance(x, Fraction):
        return x.numerator, x.denominator
    p, q = getattr(x, "p", None), getattr(x, "q", None)        # sympy Rational / Integer
    if p is not None and q is not None:
        return int(p), int(q)
    if isinstance(x, np.integer):
        return int(x), 1
    raise TypeError("the GPU elimination path takes exact integer/rational entries, got %r" % type(x))


def _kind_of(items):
    kind = "int"
    for row in items:
        for x in row:
            if isinstance(x, (int, np.integer)):
                continue
            if isinstance(x, Fraction):
                if kind == "int":
                    kind = "fraction"
            else:
                kind = "sympy"
    return kind


def _wrap(kind, p, q=1):
    if kind == "sympy":
        import sympy
        return sympy.Rational(p, q)
    if kind == "fraction":
        return Fraction(p, q)
    return Fraction(p, q) if q != 1 else p


def _to_int_grid(rows):
    """Exact entries -> (int32 array, common denominator D) with array == D * rows."""
    pq = [[_pq_of(x) for x in row] for row in rows]
    D = 1
    for row in pq:
        for _, q in row:
            if q != 1:
                D = D * q // gcd(D, q)
    grid = [[p * (D // q) for p, q in row] for row in pq]
    for row in grid:
        for v in row:
            if abs(v) > _INT32_MAX:
                raise OverflowError("entry %d does not fit the device's int32 input after clearing denominators" % v)
    return np.array(grid, dtype=np.int32).reshape(len(rows), len(rows[0]) if rows else 0), D


def _lowest(num_words, den_words):
    """Numerators ``[..., L]`` over one denominator ``[L]`` (packed device results) -> nested lists of lowest-terms
    ``(p, q)`` pairs with ``q > 0``, reduced on the device (``lsx_lowest_terms``: the reference returns reduced
    rationals, linalg.py:574, 698-699)."""
    p, q = default_engine().lowest_terms(np.ascontiguousarray(num_words)[None], np.ascontiguousarray(den_words)[None])
    P, Q = limbs_to_ints(p[0]), limbs_to_ints(q[0])

    def pairs(a, b):
        return [pairs(x, y) for x, y in zip(a, b)] if isinstance(a, list) else (a, b)
    return pairs(P, Q)


class Matrix:
    items: List[List[Any]]

    def __init__(self, items: List[List[Any]]):
        # validation as reference linalg.py:14-32
        if not items:
            raise ValueError("Matrix cannot be empty")
        if not all(isinstance(row, list) for row in items):
            raise ValueError("Matrix items must be a list of lists")
        if not items[0]:
            if any(row for row in items):
                raise ValueError("Matrix rows cannot be empty if columns exist")
            row_len = 0
        else:
            row_len = len(items[0])
            if not all(len(row) == row_len for row in items):
                raise ValueError("All matrix rows must have the same length")
        self._cols = row_len
        self.items = items

    def __str__(self) -> str:
        return "\n".join([" ".join([str(item) for item in row]) for row in self.items])

    def __repr__(self) -> str:
        return "Matrix(%r)" % (self.items,)

    def __eq__(self, other):
        return isinstance(other, Matrix) and self.items == other.items

    @property
    def rows(self) -> int:
        return len(self.items)

    @property
    def cols(self) -> int:
        if self.rows == 0:
            return self._cols
        return len(self.items[0])

    def self_map(self, f: Callable[[Any], Any]) -> "Matrix":
        return Matrix([[f(item) for item in row] for row in self.items])

    def get_row(self, i: int) -> List[Any]:
        return self.items[i]

    def get_col(self, j: int) -> List[Any]:
        return [row[j] for row in self.items]

    def inorder_slot_iter(self) -> Iterator[Tuple[int, int]]:
        for i in range(self.rows):
            for j in range(self.cols):
                yield (i, j)

    @classmethod
    def zero(cls, rows: int, cols: int) -> "Matrix":
        return cls([[0] * cols for _ in range(rows)])

    @classmethod
    def identity(cls, size: int) -> "Matrix":
        return cls([[1 if i == j else 0 for j in range(size)] for i in range(size)])

    @classmethod
    def diagonal(cls, items: List[Any]) -> "Matrix":
        res = cls.zero(len(items), len(items))
        for i, item in enumerate(items):
            res.items[i][i] = item
        return res

    @classmethod
    def new_vector(cls, items: List[Any]) -> "Matrix":
        return cls([[i] for i in items])

    def transpose(self) -> "Matrix":
        return Matrix([[self.items[j][i] for j in range(self.rows)] for i in range(self.cols)])

    def scalar_mul(self, scalar: Any) -> "Matrix":
        return Matrix([[scalar * item for item in row] for row in self.items])

    def __mul__(self, other) -> "Matrix":
        """Matrix product / scalar multiple (reference linalg.py:101-181 without its LaTeX log); host
        arithmetic on the entry objects, used by the input builder (random_matrix.py:129)."""
        if not isinstance(other, Matrix):
            return self.scalar_mul(other)
        if self.cols != other.rows:
            raise ValueError("Matrix dimensions must match")
        cols = list(zip(*other.items)) if other.items and other.items[0] else []
        return Matrix([[sum((a * b for a, b in zip(row, col)), 0) for col in cols] for row in self.items])

    class AffineSubspace:
        """Reference linalg.py:491-522 (LaTeX ``cformat`` omitted: logging is bypassed)."""

        def __init__(self, vec: List[Any], mat: "Matrix"):
            self.vec = vec
            self.generators = mat

        def get_one(self) -> List[Any]:
            return self.vec

        def dim(self) -> int:
            return self.generators.cols

        def basis(self) -> List[List[Any]]:
            return self.generators.transpose().items

        def __repr__(self):
            return "AffineSubspace(vec=%r, generators=%r)" % (self.vec, self.generators)

    class NoSolution:
        """Reference linalg.py:524-532."""

        def __init__(self):
            pass

        def __repr__(self):
            return "NoSolution()"

    # ---- the elimination path -----------------------------------------------------------------
    # step descriptions of the reference's log (linalg.py:556-559, 581, 603, 627)
    _STEP_TEXT = {1: ("S", r"Výměna řádků $R_{%d}$ a $R_{%d}$"), 2: ("N", r"Normalizace pivotního řádku %s"),
                  3: ("E", r"Eliminace prvků pod pivotem ve sloupci %s"), 4: ("E", r"Eliminace nad pivotem ve sloupci %s")}

    def row_reduce(self, bar_col: int = None, trace: bool = False):
        """Reduced row echelon form on the columns left of ``bar_col`` (reference linalg.py:534-630).

        Returns ``(A, pivots, intermediate_matrices, intermediate_steps)``.  ``bar_col`` falsy means ``cols - 1``
        exactly as linalg.py:543.  By default the two log lists are empty (the batched device path does not keep
        intermediate states).  With ``trace=True`` (integer or rational entries, rows * cols <= 4096) the device
        replays the reference's operation order and the lists are filled like the reference fills them:
        ``intermediate_steps`` holds the same ``(label, description)`` pairs (S / N / E + running number,
        linalg.py:556-561, 580-582, 601-605, 626-628) and ``intermediate_matrices`` holds the input followed by
        the matrix after every step -- as exact matrices (lists of rows), not as the LaTeX strings the reference
        renders from them: the LaTeX ``Logger`` stays outside the device path.
        """
        m, n = self.rows, self.cols
        bar = bar_col or n - 1
        kind = _kind_of(self.items)
        if trace and n > 0 and bar > 0:
            return self._row_reduce_trace(min(bar, n), kind)
        if n == 0 or bar <= 0:
            # nothing to pivot on: the loop of linalg.py:547 never runs
            return [list(r) for r in self.items], [], [], []
        if bar > n:
            bar = n
        grid, D = _to_int_grid(self.items)
        eng = default_engine()
        amax = int(np.abs(grid.astype(np.int64)).max()) if grid.size else 0
        res = eng.rref_batch(grid[None], bar, a_abs_max=amax, b_abs_max=amax)
        _raise_on_status(int(res.status[0]))
        rank = int(res.rank[0])
        pivots = [(k, int(res.pivot_col[0][k])) for k in range(rank)]
        if D == 1:                                              # integer input: lowest terms straight from the device
            return [[_wrap(kind, p, q) for p, q in row] for row in _lowest(res.num[0], res.den[0])], pivots, [], []
        num = limbs_to_ints(res.num[0])
        den = limbs_to_ints(res.den[0])
        out = []
        for i in range(m):
            # rows that never became a pivot row keep the scale D of the cleared denominators
            d_i = den if i < rank else den * D
            out.append([_wrap(kind, *reduce_pq(x, d_i)) for x in num[i]])
        return out, pivots, [], []

    def _row_reduce_trace(self, bar, kind):
        # rational entries: integer numerators over the common denominator D; the device replays the steps on the
        # residues of grid / D, so "pivot == 1" is tested on the rational entry as the reference does (linalg.py:570)
        grid, D = _to_int_grid(self.items)
        frames, ops, pivots = default_engine().rref_trace(grid, bar, den=D)
        wrap = lambda f: [[_wrap(kind, x.numerator, x.denominator) for x in row] for row in f]
        start = [[_wrap(kind, *_pq_of(x)) for x in row] for row in self.items]
        mats = [start] + [wrap(f) for f in frames]
        steps = []
        for t, (k, a, b) in enumerate(ops):
            letter, text = Matrix._STEP_TEXT[k]
            steps.append(("%s%d" % (letter, t), text % ((a + 1, b + 1) if k == 1 else (a + 1,))))
        return mats[-1], pivots, mats, steps

    def rank(self) -> int:
        """Reference linalg.py:745-747."""
        if self.cols == 0:
            return 0
        grid, _ = _to_int_grid(self.items)
        from .engine import LsxError
        try:
            res = default_engine().rank_batch(grid[None])
        except LsxError as e:
            if e.code not in (_lib.ERR_UNSUPPORTED, _lib.ERR_BOUND):
                raise
            # beyond the batched kernels (more than 254 rows, or entries that need more than 32 primes): the
            # global-memory route has no size limit, only a time one -- like the reference's own rank()
            return default_engine().rank_large(grid)[0]
        _raise_on_status(int(res.status[0]))
        return int(res.rank[0])

    def determinant(self, log_permutation_details: bool = False, use_optimal: bool = True) -> Any:
        """Reference linalg.py:183-262.  n = 0 -> 1, n = 1 -> the entry, non-square -> ValueError
        (determinant.py:772-773); otherwise sign * product of elimination pivots, from the device."""
        n = self.rows
        if n == 0:
            return 1
        if n == 1:
            return self.items[0][0]                               # the reference's shortcut for ANY one-row matrix (linalg.py:200)
        if self.rows != self.cols:
            raise ValueError("Determinant requires a square matrix")
        kind = _kind_of(self.items)
        grid, D = _to_int_grid(self.items)
        if n > 64:
            # one large matrix: residues modulo many primes (tile kernel up to n ~ 220, blocked tensor-core LU above),
            # then one CRT.  All primes run on this process's GPU: a method of one object is not a collective
            # (the by-prime multi-GPU route is linalg_solver_b200.dist.det_large_sharded, called by every rank).
            from .dist import det_large_sharded
            words, _ = det_large_sharded(default_engine(), grid, sharded=False)
            d = limbs_to_ints(np.asarray(words).reshape(1, -1))[0]
        else:
            from .engine import LsxError
            try:
                res = default_engine().det_batch(grid[None])
                _raise_on_status(int(res.status[0]))
                d = limbs_to_ints(res.det[0])
            except LsxError as e:
                if e.code != _lib.ERR_BOUND:
                    raise
                # entries so large that the batched kernels' 32 primes do not cover the Hadamard bound: the by-prime
                # route has the whole table (2048 primes)
                from .dist import det_large_sharded
                words, _ = det_large_sharded(default_engine(), grid, sharded=False)
                d = limbs_to_ints(np.asarray(words).reshape(1, -1))[0]
        return _wrap(kind, *reduce_pq(d, D ** n))

    def inverse(self, log_matrices: bool = False, log_steps: bool = False, log_result: bool = False):
        """Reference linalg.py:682-743: the inverse as a ``Matrix`` or ``Matrix.NoSolution()``.

        Without log flags the reference answers through sympy and returns sympy numbers
        (linalg.py:696-701); with a log flag it returns numbers of the input type (739-743).
        The flags only select that type here; nothing is logged."""
        if self.rows != self.cols:
            raise ValueError("Matrix must be square to invert.")
        n = self.rows
        kind = "sympy" if not (log_matrices or log_steps or log_result) else _kind_of(self.items)
        try:
            grid, D = _to_int_grid(self.items)
        except TypeError:
            if kind == "sympy":
                return Matrix.NoSolution()          # the reference swallows every exception here (700-701)
            raise
        res = default_engine().inverse_batch(grid[None])
        st = int(res.status[0])
        if st & _lib.ST_SINGULAR:
            return Matrix.NoSolution()
        _raise_on_status(st)
        if D == 1:                                              # integer input: lowest terms straight from the device
            return Matrix([[_wrap(kind, p, q) for p, q in row] for row in _lowest(res.adj[0], res.det[0])])
        adj = limbs_to_ints(res.adj[0])
        det = limbs_to_ints(res.det[0])
        # (D A)^-1 = adj / det  =>  A^-1 = D adj / det  (fractional input: the extra factor D is folded in on the host)
        return Matrix([[_wrap(kind, *reduce_pq(D * adj[i][j], det)) for j in range(n)] for i in range(n)])

    def find_preimage_of(self, vec: List[Any], log_matrices: bool = False, log_steps: bool = False,
                         log_result: bool = False):
        """Solution set of ``self * x = vec`` (reference linalg.py:632-680).

        Default route (no log flag) mirrors ``_q_find_preimage_of`` (870-910): sympy numbers, generator
        columns ordered like ``sorted(params, key=str)`` (895), ``Matrix.zero(n, 0)`` generators for a
        unique solution (888).  With a log flag it mirrors ``_extract_affine_subspace`` (937-999):
        input-typed numbers, ascending free-column order, ``None`` generators for a unique solution."""
        if self.rows != len(vec):
            raise ValueError("Matrix dimensions must match")
        logged = log_matrices or log_steps or log_result
        m, n = self.rows, self.cols
        aug = [list(self.items[i]) + [vec[i]] for i in range(m)]
        kind = _kind_of(aug) if logged else "sympy"
        grid, _ = _to_int_grid(aug)
        A = np.ascontiguousarray(grid[:, :n])
        b = np.ascontiguousarray(grid[:, n])
        res = default_engine().solve_batch(A[None], b[None], gen_cap=n)
        st = int(res.status[0])
        if st & _lib.ST_INCONSISTENT:
            return Matrix.NoSolution()
        _raise_on_status(st)
        rank = int(res.rank[0])
        k = n - rank
        # x = (D A)^-1 (D b): the common denominator of the inputs cancels, so the device's lowest terms are final
        particular = [_wrap(kind, p, q) for p, q in _lowest(res.particular[0], res.den[0])]
        if k == 0:
            return Matrix.AffineSubspace(particular, None if logged else Matrix.zero(n, 0))
        gens = _lowest(res.generators[0], res.den[0])    # [n][gen_cap] pairs
        order = list(range(k)) if logged else sorted(range(k), key=lambda i: "tau%d" % i)
        gen_items = [[_wrap(kind, *gens[r][c]) for c in order] for r in range(n)]
        return Matrix.AffineSubspace(particular, Matrix(gen_items))

    def kernel(self):
        """Reference linalg.py:749-756."""
        return self.find_preimage_of([0] * self.rows)

    def find_eigenspace(self, eigenvalue: Any):
        """Reference linalg.py:758-770: nullspace of ``A - eigenvalue * I`` through ``kernel()`` (SURVEY.md
        section 8f item 1; the eigenvalue must be an exact rational for the device path)."""
        if self.rows != self.cols:
            raise ValueError("Matrix must be square to find eigenspace.")
        items = [list(row) for row in self.items]
        for i in range(self.rows):
            items[i][i] = items[i][i] - eigenvalue
        return Matrix(items).kernel()


    # ---- eigen-callers on the device engine --------------------
    def characteristic_polynomial(self) -> List[Any]:
        """Coefficients ``[c_0, ..., c_n]`` of ``det(A - lambda I) = sum c_k lambda^k`` (exact rationals).

        The reference expands this determinant over ``Polynomial`` entries through its planner
        (linalg.py:424-445, determinant.py:761-803).  Here it is n + 1 integer determinants
        ``det(A - t I)``, t = 0..n, in ONE batched device call (``lsx_det_batch``), followed by exact Newton
        interpolation on the host: a polynomial of degree n is fixed by n + 1 values.
        """
        if self.rows != self.cols:
            raise ValueError("Eigenvalues require a square matrix")
        n = self.rows
        kind = _kind_of(self.items)
        grid, D = _to_int_grid(self.items)                       # grid = D * A
        pts = list(range(n + 1))
        batch = np.repeat(grid[None].astype(np.int64), n + 1, axis=0)
        for k, t in enumerate(pts):
            batch[k, range(n), range(n)] -= D * t                # D * (A - t I)
        if np.abs(batch).max(initial=0) > _INT32_MAX:
            raise OverflowError("entries of A - t I do not fit the device's int32 input")
        res = default_engine().det_batch(batch.astype(np.int32))
        for st in res.status:
            _raise_on_status(int(st))
        vals = [Fraction(v, D ** n) for v in limbs_to_ints(res.det)]          # det(A - t I)
        # Newton divided differences on the nodes 0..n, then expansion to monomial coefficients
        dd = list(vals)
        for j in range(1, n + 1):
            for i in range(n, j - 1, -1):
                dd[i] = (dd[i] - dd[i - 1]) / (pts[i] - pts[i - j])
        coeffs = [Fraction(0)] * (n + 1)
        for i in range(n, -1, -1):                               # Horner: coeffs = coeffs * (x - pts[i]) + dd[i]
            nxt = [Fraction(0)] * (n + 1)
            for k in range(n):
                nxt[k + 1] += coeffs[k]
                nxt[k] -= coeffs[k] * pts[i]
            nxt[0] += dd[i]
            coeffs = nxt
        return [_wrap("sympy" if kind == "sympy" else "fraction", c.numerator, c.denominator) for c in coeffs]

    def eigenvalues(self, real_only: bool = False):
        """``{root: algebraic multiplicity}`` of the characteristic polynomial (reference linalg.py:424-480; its
        LaTeX log is bypassed).  Root finding is sympy's, as the reference's ``radical_roots`` ends there too."""
        import sympy
        lam = sympy.Symbol("lambda")
        coeffs = self.characteristic_polynomial()
        poly = sympy.Poly(sum(sympy.Rational(*_pq_of(c)) * lam ** k for k, c in enumerate(coeffs)), lam)
        roots = sympy.roots(poly)
        if real_only:
            roots = {r: m for r, m in roots.items() if r.is_real is True}
        return roots

    def eigenvalues_with_geometric_multiplicities(self):
        """Reference linalg.py:808-818: eigenvalue -> (algebraic, geometric multiplicity); the geometric one is
        the dimension of ``find_eigenspace`` (device ``kernel``) for rational eigenvalues."""
        result = {}
        for eig, alg in self.eigenvalues().items():
            try:
                space = self.find_eigenspace(eig)
                geom = space.dim() if hasattr(space, "dim") else 0
            except TypeError:                                    # irrational eigenvalue: not on the device path
                geom = None
            result[eig] = (alg, geom)
        return result

    class DiagonalizationResult:
        """Reference linalg.py:772-806 (``cformat`` omitted: logging is bypassed)."""

        def __init__(self, eig_mults, success, P=None, P_inv=None, D=None):
            self.eigenvalue_multiplicities = eig_mults
            self.success = success
            self.P, self.P_inv, self.D = P, P_inv, D

        def __repr__(self):
            return ("DiagonalizationResult(success=%s, eigenvalue_multiplicities=%s, P=%s, P_inv=%s, D=%s)"
                    % (self.success, self.eigenvalue_multiplicities, self.P, self.P_inv, self.D))

    def diagonalize(self):
        """Reference linalg.py:833-863: eigenspaces through ``kernel`` and ``P.inverse()`` on the device."""
        if self.rows != self.cols:
            raise ValueError("Matrix must be square to diagonalize.")
        n = self.rows
        eig_mults = self.eigenvalues_with_geometric_multiplicities()
        basis_vectors = []
        for eig, (_, geom) in eig_mults.items():
            if geom is None:
                return Matrix.DiagonalizationResult(eig_mults, False)
            space = self.find_eigenspace(eig)
            if hasattr(space, "basis"):
                basis_vectors.extend(space.basis())
        if len(basis_vectors) != n:
            return Matrix.DiagonalizationResult(eig_mults, False)
        P = Matrix([list(col) for col in zip(*basis_vectors)])
        P_inv = P.inverse()
        if isinstance(P_inv, Matrix.NoSolution):
            return Matrix.DiagonalizationResult(eig_mults, False)
        return Matrix.DiagonalizationResult(eig_mults, True, P, P_inv, P_inv * self * P)


def _raise_on_status(st):
    if st & _lib.ST_BOUND:
        raise OverflowError("liblsx: an entry exceeded the declared magnitude bound")
    if st & _lib.ST_NO_GOOD_PRIME:
        raise RuntimeError("liblsx: could not collect enough agreeing primes")
    if st & _lib.ST_GEN_TRUNC:
        raise RuntimeError("liblsx: generator buffer too small")
