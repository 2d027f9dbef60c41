"""``RandomMatrixBuilder`` with the reference's builder surface for the input distributions of the
elimination path (reference linalg_solver/random_matrix.py:7-129), its rejection loops served by the
device ``rank()``.

Same attribute names, chaining methods and draw order as the reference (``new``, ``with_size``,
``with_rank``, ``with_dist``, ``build_sized``, ``build``, ``build_random``, ``build_full_rank``,
``build_rank``), so that under the same ``random.seed`` it yields the same matrices: entries come from
``dist()`` (default ``random.randint(-5, 5)``, random_matrix.py:104) row by row; ``build_full_rank``
redraws until ``rank() == N`` (109-115); ``build_rank`` is the product of a full-rank ``rows x r`` and a
full-rank ``r x cols`` factor (117-129).  The reference pays a sympy rank per candidate here, which makes
``with_rank`` unusable beyond a few dozen rows; the device rank is exact and takes
microseconds.  Eigenvalue / Jordan-block builders (131-167, 233-267) are outside the elimination path
and are not provided.
"""
import random
from typing import Any, Callable, Optional

from .matrix import Matrix


class RandomMatrixBuilder:
    rank: Optional[int] = None
    num_rows: Optional[int] = None
    num_cols: Optional[int] = None
    dist: Optional[Callable[[], Any]] = None

    @classmethod
    def new(cls, **kwargs) -> "RandomMatrixBuilder":
        builder = cls()
        for key, value in kwargs.items():
            setattr(builder, key, value)
        return builder

    def with_size(self, num_rows: int, num_cols: int) -> "RandomMatrixBuilder":
        self.num_rows, self.num_cols = num_rows, num_cols
        return self

    def with_rank(self, rank: int) -> "RandomMatrixBuilder":
        self.rank = rank
        return self

    def with_dist(self, dist: Callable[[], Any]) -> "RandomMatrixBuilder":
        self.dist = dist
        return self

    def is_square(self) -> bool:
        return self.num_rows == self.num_cols

    def assert_requirements(self) -> None:
        if self.rank is not None:
            assert self.rank <= min(self.num_rows, self.num_cols), "Rank cannot exceed min(num_rows, num_cols)."

    def build_sized(self, num_rows: int, num_cols: Optional[int] = None) -> Matrix:
        self.num_rows = num_rows
        self.num_cols = num_cols if num_cols is not None else num_rows
        return self.build()

    def build(self) -> Matrix:
        self.assert_requirements()
        if self.rank is not None:
            if self.rank == min(self.num_rows, self.num_cols) and self.num_rows == self.num_cols:
                return self.build_full_rank()
            return self.build_rank()
        return self.build_random()

    def _draw(self, rows: int, cols: int) -> Matrix:
        dist = self.dist or (lambda: random.randint(-5, 5))
        return Matrix([[dist() for _ in range(cols)] for _ in range(rows)])

    def build_random(self) -> Matrix:
        return self._draw(self.num_rows, self.num_cols)

    def build_full_rank(self) -> Matrix:
        n = self.num_rows
        while True:
            val = self._draw(n, n)
            if val.rank() == n:
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
