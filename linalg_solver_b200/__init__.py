"""B200-native exact elimination engine behind the ``Matrix`` API of koskja/linalg-solver.

Only the hot path is here: ``Matrix.row_reduce`` and the ``determinant``,
``inverse``, ``rank``, ``kernel`` and ``find_preimage_of`` calls that sit on it, plus the batched
entry points the C-ABI adds.  Importing this package loads ``liblsx.so`` (hand-written sm_100a
CUDA); it raises if the library has not been built -- there is no CPU fallback.
"""
from . import _lib                                   # noqa: F401  (fails loudly when liblsx.so is missing)
from .engine import (DetResult, Engine, InverseResult, LsxError, RankResult, RrefResult,   # noqa: F401
                     SolveResult, default_engine)
from .matrix import Matrix                            # noqa: F401
from .random_matrix import RandomMatrixBuilder        # noqa: F401
from . import convert, dist                           # noqa: F401

AffineSubspace = Matrix.AffineSubspace
NoSolution = Matrix.NoSolution

__all__ = ["Matrix", "AffineSubspace", "NoSolution", "RandomMatrixBuilder", "Engine", "LsxError", "default_engine", "convert", "dist"]
