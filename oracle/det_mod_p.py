"""TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's cpu_baseline leg, never by the product).

det(A) mod p by plain Gaussian elimination over Z/p in numpy int64: the independent CPU check SURVEY.md
section 8(c) asks for where the reference itself is infeasible (config 5, single large determinant).  It
follows the forward sweep of reference linalg.py:547-609 (pivot = entry at the diagonal if non-zero, else the
first lower non-zero row; row_i -= factor * pivot_row for the rows below), tracking sign * product of pivots,
with every operation reduced modulo p.  Products of two residues below 2^31 fit int64, so it is exact.
Pinned by tests/test_oracle_golden.py against the DomainMatrix determinants of tests/golden/c5_standins.
"""
import numpy as np


def det_mod_p(A, p: int) -> int:
    p = int(p)
    W = np.mod(np.asarray(A, dtype=np.int64), p)
    n = W.shape[0]
    if W.shape != (n, n):
        raise ValueError("Determinant requires a square matrix")
    det = 1
    for j in range(n):
        col = W[j:, j]
        nz = np.flatnonzero(col)
        if nz.size == 0:
            return 0
        src = j + int(nz[0])
        if src != j:
            W[[j, src]] = W[[src, j]]
            det = -det
        piv = int(W[j, j])
        det = det * piv % p
        if j + 1 < n:
            inv = pow(piv, p - 2, p)
            f = W[j + 1:, j] * inv % p                       # multipliers, < p
            W[j + 1:, j + 1:] = (W[j + 1:, j + 1:] - np.outer(f, W[j, j + 1:]) % p) % p
    return det % p
