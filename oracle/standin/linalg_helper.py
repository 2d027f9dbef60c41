"""Pure-Python stand-in for the reference's Rust module ``linalg_helper``.

TEST INFRASTRUCTURE ONLY.  The reference's ``linalg_solver`` package hard-imports
the pyo3 module ``linalg_helper`` (reference linalg_solver/permutation.py:1,
determinant.py:11).  That crate cannot be built in this image (no cargo/rustc),
so ``oracle/gen_golden.py`` puts this directory on ``sys.path`` to let the
UNMODIFIED reference Python import.  Only what the elimination path and the
legacy determinant touch is provided:

* ``Permutation(list)``, ``__call__``, ``sign`` -- behaviour of
  linalg-helper/src/permutation.rs:107-132 and 176-187 (sign = (-1)^(n - #cycles)).
* ``RowColPermutation`` and ``find_optimal_determinant_process`` exist as names
  only; the Rust-planned determinant (``use_optimal=True``) is not runnable here
  and the golden generator never calls it.

Nothing in the product package imports this file.
"""


class Permutation:
    def __init__(self, perm):
        perm = list(perm)
        n = len(perm)
        seen = [False] * n
        for p in perm:
            if not (0 <= p < n) or seen[p]:
                raise ValueError("Input list is not a valid permutation of 0..n-1")
            seen[p] = True
        self.perm = perm

    def __call__(self, i):
        return self.perm[i]

    def sign(self):
        n = len(self.perm)
        if n == 0:
            return 1
        seen = [False] * n
        cycles = 0
        for i in range(n):
            if not seen[i]:
                cycles += 1
                j = i
                while not seen[j]:
                    seen[j] = True
                    j = self.perm[j]
        return 1 if (n - cycles) % 2 == 0 else -1

    def cformat(self, arg_of=""):
        return str(self.perm)


class RowColPermutation:
    def __init__(self, *a, **k):
        raise NotImplementedError("stand-in: the Rust planner is not available")


def find_optimal_determinant_process(*a, **k):
    raise NotImplementedError("stand-in: the Rust planner is not available")
