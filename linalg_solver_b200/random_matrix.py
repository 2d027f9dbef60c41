"""``RandomMatrixBuilder`` and the ``gen_*`` helpers: the reference's input generators
(reference linalg_solver/random_matrix.py:7-267) on the device engine.

This module keeps the reference's builder surface on purpose -- attribute names, chaining methods and, above
all, the ORDER in which random numbers are drawn -- because the parity tests feed the reference and this
package from the same ``random.seed`` and need the same matrices out of both (it is a re-statement of the
builder's control flow, not new design; the arithmetic behind it is what differs):

* ``build_random``: entries from ``dist()`` (default ``random.randint(-5, 5)``) row by row (103-107);
* ``build_full_rank`` / ``build_rank``: rejection loops on ``Matrix.rank()`` (109-129) -- one device rank per
  candidate instead of a sympy rank, which is what makes ``with_rank`` usable beyond a few dozen rows
  (SURVEY.md section 6);
* ``build_diagonalizable`` / ``build_jordanized``: ``P^-1 D P`` and ``P^-1 J P`` with ``P`` from
  ``gen_unimodular_matrix`` (131-167, 233-267) -- the ``P.inverse()`` is the device inverse (SURVEY.md section
  8f-2); the products are host arithmetic on the entry objects, as in the reference.
"""
import random
from typing import Any, Callable, List, Optional, Tuple

from .matrix import Matrix


def _default_dist() -> int:
    return random.randint(-5, 5)


class RandomMatrixBuilder:
    rank: Optional[int] = None
    eigenvalues: Optional[List[Tuple[Any, int]]] = None
    jordan_blocks: Optional[List[Tuple[Any, int]]] = None
    do_randomize_from_diagonal_form: bool = True
    num_rows: Optional[int] = None
    num_cols: Optional[int] = None
    dist: Optional[Callable[[], Any]] = None

    @classmethod
    def new(cls, **kwargs) -> "RandomMatrixBuilder":
        builder = cls()
        for key, value in kwargs.items():
            setattr(builder, key, value)
        return builder

    # ---- chaining setters -----------------------------------------------------------------
    def with_size(self, num_rows: int, num_cols: int) -> "RandomMatrixBuilder":
        self.num_rows, self.num_cols = num_rows, num_cols
        return self

    def with_rank(self, rank: int) -> "RandomMatrixBuilder":
        self.rank = rank
        return self

    def with_dist(self, dist: Callable[[], Any]) -> "RandomMatrixBuilder":
        self.dist = dist
        return self

    def with_eigenvalues(self, eigenvalues) -> "RandomMatrixBuilder":
        """Eigenvalues as ``(value, multiplicity)`` pairs or bare values (multiplicity 1), random_matrix.py:36-43."""
        pairs = isinstance(eigenvalues[0], tuple)
        self.eigenvalues = list(eigenvalues) if pairs else [(e, 1) for e in eigenvalues]
        return self

    def with_jordan_blocks(self, blocks: List[Tuple[Any, int]]) -> "RandomMatrixBuilder":
        self.jordan_blocks = blocks
        return self

    def is_square(self) -> bool:
        return self.num_rows == self.num_cols

    def assert_requirements(self) -> None:
        """The reference's consistency checks (random_matrix.py:54-80), same messages."""
        spec = {"eigenvalues": self.eigenvalues is not None, "rank": self.rank is not None,
                "Jordan blocks": self.jordan_blocks is not None}
        if spec["eigenvalues"]:
            assert self.is_square(), "Diagonalizable matrix must be square."
            assert sum(e[1] for e in self.eigenvalues) == self.num_rows, \
                "Sum of eigenvalue multiplicities must match matrix size."
            assert not spec["rank"], "Cannot specify both eigenvalues and rank."
            assert not spec["Jordan blocks"], "Cannot specify both eigenvalues and Jordan blocks."
        if spec["rank"]:
            assert self.rank <= min(self.num_rows, self.num_cols), "Rank cannot exceed min(num_rows, num_cols)."
            assert not spec["eigenvalues"], "Cannot specify both rank and eigenvalues."
            assert not spec["Jordan blocks"], "Cannot specify both rank and Jordan blocks."
        if spec["Jordan blocks"]:
            assert self.is_square(), "Jordan block matrix must be square."
            assert sum(size for _, size in self.jordan_blocks) == self.num_rows, \
                "Sum of Jordan block sizes must match matrix size."
            assert not spec["eigenvalues"], "Cannot specify both Jordan blocks and eigenvalues."
            assert not spec["rank"], "Cannot specify both Jordan blocks and rank."

    # ---- builders -------------------------------------------------------------------------
    def build_sized(self, num_rows: int, num_cols: Optional[int] = None) -> Matrix:
        self.num_rows = num_rows
        self.num_cols = num_cols if num_cols is not None else num_rows
        return self.build()

    def build(self) -> Matrix:
        """Dispatch of random_matrix.py:87-101: Jordan blocks, then eigenvalues, then rank, else plain random."""
        self.assert_requirements()
        if self.jordan_blocks is not None:
            return self.build_jordanized()
        if self.eigenvalues is not None:
            return self.build_diagonalizable()
        if self.rank is not None:
            if self.rank == min(self.num_rows, self.num_cols) and self.num_rows == self.num_cols:
                return self.build_full_rank()
            return self.build_rank()
        return self.build_random()

    def _draw(self, rows: int, cols: int) -> Matrix:
        dist = self.dist or _default_dist
        return Matrix([[dist() for _ in range(cols)] for _ in range(rows)])

    def build_random(self) -> Matrix:
        return self._draw(self.num_rows, self.num_cols)

    def build_full_rank(self) -> Matrix:
        n = self.num_rows
        while True:
            val = self._draw(n, n)
            if val.rank() == n:                                   # device rank (reference: sympy rank, 109-115)
                return val

    def build_rank(self) -> Matrix:
        rows, cols, rank = self.num_rows, self.num_cols, self.rank
        while True:
            a = self._draw(rows, rank)
            if a.rank() == rank:
                break
        while True:
            b = self._draw(rank, cols)
            if b.rank() == rank:
                break
        return a * b

    def build_diagonalizable(self) -> Matrix:
        """``P^-1 D P`` for the diagonal ``D`` of the requested eigenvalues (random_matrix.py:131-142)."""
        diag = [eig for eig, mult in self.eigenvalues for _ in range(mult)]
        D = Matrix.diagonal(diag)
        if not self.do_randomize_from_diagonal_form:
            return D
        P = gen_unimodular_matrix(self.num_rows)
        return P.inverse() * D * P                                # device inverse; exact integers (det P = +-1)

    def build_jordan(self) -> Matrix:
        """Block-diagonal Jordan form of ``jordan_blocks`` (random_matrix.py:144-158)."""
        n = self.num_rows
        total = sum(size for _, size in self.jordan_blocks)
        if total != n:
            raise ValueError(f"Sum of Jordan block sizes ({total}) must equal matrix size ({n})")
        J = [[0] * n for _ in range(n)]
        at = 0
        for eigenvalue, size in self.jordan_blocks:
            for i in range(size):
                J[at + i][at + i] = eigenvalue
                if i + 1 < size:
                    J[at + i][at + i + 1] = 1
            at += size
        return Matrix(J)

    def build_jordanized(self) -> Matrix:
        """``P^-1 J P`` for a random unimodular ``P`` (random_matrix.py:160-167)."""
        J = self.build_jordan()
        P = gen_unimodular_matrix(self.num_rows)
        return P.inverse() * J * P


# ---- module-level helpers of the reference (random_matrix.py:170-267) ------------------------
def raw_gen_rand_matrix(rows: int, cols: int, dist: Optional[Callable[[], Any]] = None) -> Matrix:
    return RandomMatrixBuilder.new().with_size(rows, cols).with_dist(dist).build_random()


def gen_regular_matrix(N: int, dist: Optional[Callable[[], Any]] = None) -> Matrix:
    return RandomMatrixBuilder.new().with_size(N, N).with_dist(dist).build_full_rank()


def gen_matrix_with_rank(rows: int, cols: int, rank: Optional[int] = None,
                         dist: Optional[Callable[[], Any]] = None) -> Matrix:
    return RandomMatrixBuilder.new().with_size(rows, cols).with_rank(rank or min(rows, cols)).with_dist(dist).build_rank()


def gen_jordan_matrix(N: int, blocks: List[Tuple[Any, int]]) -> Matrix:
    return RandomMatrixBuilder.new().with_size(N, N).with_jordan_blocks(blocks).build_jordan()


def gen_matrix_with_jordan_blocks(N: int, blocks: List[Tuple[Any, int]],
                                  dist: Optional[Callable[[], Any]] = None) -> Matrix:
    return RandomMatrixBuilder.new().with_size(N, N).with_jordan_blocks(blocks).with_dist(dist).build_jordanized()


def gen_diagonalizable_matrix(N: int, eigenvalues: Optional[List[Tuple[Any, int]]] = None,
                              dist: Optional[Callable[[], Any]] = None) -> Matrix:
    if eigenvalues is None:
        eigenvalues = [(dist() if dist is not None else random.randint(-5, 5), 1) for _ in range(N)]
    return RandomMatrixBuilder.new().with_size(N, N).with_eigenvalues(eigenvalues).with_dist(dist).build_diagonalizable()


def gen_unimodular_matrix(N: int, dist: Optional[Callable[[], Any]] = None) -> Matrix:
    """``L * U`` of a lower and an upper triangular matrix with +-1 on the diagonal, so det = +-1
    (random_matrix.py:233-267).  Draw order as in the reference: U row by row (sign of the diagonal entry, then the
    entries right of it), then L row by row (sign, then the entries left of it); off-diagonal default
    ``random.randint(-1, 1)``."""
    if dist is None:
        dist = lambda: random.randint(-1, 1)
    sign = lambda: random.choice([-1, 1])
    U = [[0] * N for _ in range(N)]
    for i in range(N):
        U[i][i] = sign()
        for j in range(i + 1, N):
            U[i][j] = dist()
    L = [[0] * N for _ in range(N)]
    for i in range(N):
        L[i][i] = sign()
        for j in range(i):
            L[i][j] = dist()
    return Matrix(L) * Matrix(U)
