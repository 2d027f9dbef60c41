"""The reference-facing ``Matrix`` API served by the device: same calls, same return shapes and types
as reference linalg.py, checked against the reference-generated golden fixtures."""
from fractions import Fraction

import pytest

from oracle import golden_io, ref_port

pytestmark = pytest.mark.gpu


def pq(x):
    if isinstance(x, int):
        return [x, 1]
    if isinstance(x, Fraction):
        return [x.numerator, x.denominator]
    return [int(x.p), int(x.q)]


def rat(items):
    import sympy
    return [[sympy.Rational(x) for x in row] for row in items]


def test_row_reduce_signature_and_values():
    from linalg_solver_b200 import Matrix
    g = golden_io.load("edge_small")
    for c in g["rref_cases"][::7]:
        M = Matrix(rat(c["A"]))
        R, piv, frames, steps = M.row_reduce(c["bar_col"]) if c["bar_col"] is not None else M.row_reduce()
        assert frames == [] and steps == []
        assert [pq(x) for row in R for x in row] == c["rref"], (c["A"], c["bar_col"])
        assert [list(p) for p in piv] == c["pivots"]
        assert all(isinstance(p, tuple) for p in piv)
        assert M.rank() == c["rank"] and isinstance(M.rank(), int)


def test_row_reduce_types_follow_input():
    import sympy
    from linalg_solver_b200 import Matrix
    A = [[2, 1, 1], [4, 3, 1]]
    R, _, _, _ = Matrix([[Fraction(x) for x in r] for r in A]).row_reduce(2)
    assert all(isinstance(x, Fraction) for row in R for x in row)
    R2, _, _, _ = Matrix(rat(A)).row_reduce(2)
    assert all(isinstance(x, sympy.Rational) for row in R2 for x in row)
    assert [pq(x) for row in R for x in row] == [pq(x) for row in R2 for x in row]
    want, _ = ref_port.row_reduce(A, 2)
    assert [pq(x) for row in R for x in row] == [[x.numerator, x.denominator] for row in want for x in row]


def test_fractional_entries_are_cleared_and_scaled_back():
    from linalg_solver_b200 import Matrix
    A = [[Fraction(1, 2), Fraction(1, 3), 1, 0], [Fraction(1, 2), Fraction(1, 3), 0, 1], [1, Fraction(-2, 5), 3, 3]]
    R, piv, _, _ = Matrix([list(r) for r in A]).row_reduce(2)
    want, wpiv = ref_port.row_reduce(A, 2)
    assert R == want and piv == wpiv
    B = [[Fraction(1, 2), Fraction(1, 3)], [Fraction(3, 4), Fraction(-1, 5)]]
    inv = Matrix([list(r) for r in B]).inverse(log_result=True)
    assert inv.items == ref_port.inverse(B)
    assert Matrix([list(r) for r in B]).determinant() == ref_port.determinant(B)


def test_inverse_default_and_logged_routes():
    import sympy
    from linalg_solver_b200 import Matrix
    g = golden_io.load("edge_small")
    for c in g["inverse_cases"]:
        M = Matrix(rat(c["A"]))
        inv = M.inverse()
        if c["default"] is None:
            assert isinstance(inv, Matrix.NoSolution) and repr(inv) == "NoSolution()"
            assert isinstance(M.inverse(log_steps=True), Matrix.NoSolution)
        else:
            assert [pq(x) for row in inv.items for x in row] == c["default"]
            assert all(isinstance(x, sympy.Rational) for row in inv.items for x in row)
            inv2 = Matrix([list(r) for r in c["A"]]).inverse(log_matrices=True)
            assert [pq(x) for row in inv2.items for x in row] == c["logged"]
        assert pq(M.determinant()) == c["det"]
    with pytest.raises(ValueError, match="Matrix must be square to invert."):
        Matrix([[1, 2, 3], [4, 5, 6]]).inverse()
    with pytest.raises(ValueError):
        Matrix([[1, 2, 3], [4, 5, 6]]).determinant()
    assert Matrix([[7]]).determinant() == 7


def test_find_preimage_and_kernel_routes():
    from linalg_solver_b200 import Matrix
    g = golden_io.load("edge_small")
    seen_nosol = seen_many = 0
    for c in g["system_cases"]:
        M = Matrix(rat(c["A"]))
        n = len(c["A"][0])
        for route, kw in (("default", {}), ("logged", {"log_result": True})):
            want = c[route]
            res = M.find_preimage_of([__import__("sympy").Rational(x) for x in c["b"]], **kw)
            if want["status"] != "ok":
                assert isinstance(res, Matrix.NoSolution)
                seen_nosol += 1
                continue
            assert isinstance(res, Matrix.AffineSubspace)
            if want["gen_cols"] is None:
                assert res.generators is None
                gens = []
            else:
                assert res.generators.cols == want["gen_cols"] and res.generators.rows == n
                assert res.dim() == want["gen_cols"]
                gens = [pq(x) for row in res.generators.items for x in row]
                seen_many += want["gen_cols"] > 10
            part = [pq(x) for x in res.vec]
            assert golden_io.digest_pq(part + gens) == want["sha"]
            if "particular" in want:
                assert part == want["particular"] and gens == want["generators"]
    assert seen_nosol > 0 and seen_many > 0
    with pytest.raises(ValueError, match="Matrix dimensions must match"):
        Matrix([[1, 2], [3, 4]]).find_preimage_of([1, 2, 3])
    ker = Matrix([[1, 2], [2, 4]]).kernel()
    assert [pq(x) for x in ker.vec] == [[0, 1], [0, 1]]
    assert [pq(x) for row in ker.generators.items for x in row] == [[-2, 1], [1, 1]]
    assert ker.basis() == ker.generators.transpose().items


def test_random_matrix_builder_reproduces_reference_inputs():
    """Same seed, same draw order, device rank() in the rejection loops (reference
    random_matrix.py:103-129): the builder must yield the reference's own matrices (golden inputs)."""
    import random
    from linalg_solver_b200 import Matrix, RandomMatrixBuilder
    g1 = golden_io.load("c1_4x4")
    random.seed(20260001)
    for c in g1["cases"][:200]:
        assert RandomMatrixBuilder.new().with_size(4, 4).build().items == c["A"]
    g3 = golden_io.load("c3_16x17")
    for i in (0, 1, 2, 7, 500):
        random.seed(202600030000 + i)
        A = RandomMatrixBuilder.new().with_size(16, 16).with_rank(10).build()
        assert A.items == g3["cases"][i]["A"]
        assert A.rank() == 10
    random.seed(5)
    F = RandomMatrixBuilder.new(num_rows=6, num_cols=6, rank=6).build()
    assert F.rank() == 6 and not isinstance(F.inverse(), Matrix.NoSolution)
    big = RandomMatrixBuilder.new().with_size(64, 64).with_rank(48).build()     # infeasible in the reference
    assert big.rank() == 48 and big.kernel().dim() == 16
    with pytest.raises(AssertionError):
        RandomMatrixBuilder.new().with_size(3, 3).with_rank(4).build()


def test_find_eigenspace_and_mul():
    import sympy
    from linalg_solver_b200 import Matrix
    A = Matrix([[2, 0, 0], [0, 3, 4], [0, 4, 9]])
    es = A.find_eigenspace(11)
    assert es.dim() == 1
    v = [x[0] for x in es.generators.items]
    assert [pq(x) for x in v] == [[0, 1], [1, 2], [1, 1]]
    assert A.find_eigenspace(sympy.Rational(5)).dim() == 0
    with pytest.raises(ValueError, match="Matrix must be square to find eigenspace."):
        Matrix([[1, 2, 3]]).find_eigenspace(1)
    P = Matrix([[1, 2], [3, 4]]) * Matrix([[0, 1], [1, 0]])
    assert P.items == [[2, 1], [4, 3]] and (Matrix([[1, 2]]) * 3).items == [[3, 6]]


def test_characteristic_polynomial_matches_sympy():
    """det(A - lambda I) from n + 1 batched device determinants + interpolation against
    sympy's exact characteristic polynomial, for integer and fractional entries."""
    import random
    import sympy
    from linalg_solver_b200 import Matrix
    rnd = random.Random(8)
    lam = sympy.Symbol("lambda")
    for n in (1, 2, 3, 5, 8, 12):
        items = [[rnd.randint(-5, 5) for _ in range(n)] for _ in range(n)]
        if n == 5:
            items = [[Fraction(x, rnd.choice([1, 2, 3])) for x in row] for row in items]
        coeffs = Matrix(items).characteristic_polynomial()
        # sympy's charpoly is det(lambda I - A); det(A - lambda I) = (-1)^n times it
        sm = sympy.Matrix([[sympy.Rational(x.numerator, x.denominator) if isinstance(x, Fraction) else x for x in row]
                           for row in items])
        want = [(-1) ** n * w for w in sm.charpoly(lam).all_coeffs()[::-1]]
        assert [pq(c) for c in coeffs] == [pq(sympy.Rational(w)) for w in want], n


def test_eigenvalues_and_diagonalize_on_the_device():
    """Reference linalg.py:424-480, 808-863: eigenvalues with multiplicities, eigenspaces through kernel(),
    P^-1 through inverse(); a diagonalizable matrix built as P D P^-1 and a Jordan block that is not."""
    import sympy
    from linalg_solver_b200 import Matrix
    P = sympy.Matrix([[1, 2, 0, 1], [0, 1, 3, 0], [1, 0, 1, 1], [2, 1, 0, 3]])
    assert P.det() != 0
    D = sympy.diag(2, 2, -1, 5)
    A = P * D * P.inv()
    M = Matrix([[sympy.Rational(A[i, j]) for j in range(4)] for i in range(4)])
    assert M.eigenvalues() == {sympy.Integer(2): 2, sympy.Integer(-1): 1, sympy.Integer(5): 1}
    mult = M.eigenvalues_with_geometric_multiplicities()
    assert mult[sympy.Integer(2)] == (2, 2) and mult[sympy.Integer(5)] == (1, 1)
    res = M.diagonalize()
    assert res.success
    Dm = res.D.items
    assert all(Dm[i][j] == 0 for i in range(4) for j in range(4) if i != j)
    assert sorted(Dm[i][i] for i in range(4)) == [-1, 2, 2, 5]
    assert (res.P_inv * res.P).items == Matrix.identity(4).items
    J = Matrix([[3, 1, 0], [0, 3, 1], [0, 0, 3]])
    assert J.eigenvalues() == {sympy.Integer(3): 3}
    rj = J.diagonalize()
    assert not rj.success and rj.eigenvalue_multiplicities[sympy.Integer(3)] == (3, 1)
    with pytest.raises(ValueError):
        Matrix([[1, 2, 3], [4, 5, 6]]).diagonalize()


def test_row_reduce_trace_matches_reference_steps_and_frames():
    """row_reduce(trace=True): step labels, descriptions and every intermediate matrix against the frames of the
    unmodified reference (tests/golden/trace_small)."""
    from linalg_solver_b200 import Matrix
    g = golden_io.load("trace_small")
    seen = set()
    for c in g["cases"]:
        M = Matrix(rat(c["A"]))
        R, piv, mats, steps = M.row_reduce(c["bar_col"], trace=True) if c["bar_col"] is not None else M.row_reduce(trace=True)
        assert [list(s) for s in steps] == c["steps"], (c["A"], c["bar_col"])
        assert [[pq(x) for row in f for x in row] for f in mats] == c["frames"]
        assert [pq(x) for row in R for x in row] == c["rref"] and [list(p) for p in piv] == c["pivots"]
        seen |= {s[0][0] for s in steps}
    assert seen == {"S", "N", "E"}
    # the default call keeps its empty log lists and agrees with the traced result
    M = Matrix(rat(g["cases"][40]["A"]))
    R0, p0, f0, s0 = M.row_reduce()
    R1, p1, f1, s1 = M.row_reduce(trace=True)
    assert f0 == [] and s0 == [] and R0 == R1 and p0 == p1 and len(f1) == len(s1) + 1


def test_determinant_of_a_large_matrix_goes_through_the_by_prime_route():
    """Matrix.determinant() beyond the batched kernels (n > 64): residues + CRT, exact against the DomainMatrix
    stand-ins (128: tile kernel per prime, 256: blocked tensor-core LU)."""
    import numpy as np
    from linalg_solver_b200 import Matrix
    g = golden_io.load("c5_standins")
    for c in g["cases"]:
        if c["n"] not in (128, 256):
            continue
        rng = np.random.Generator(np.random.PCG64(c["seed"]))
        A = rng.integers(-5, 6, size=(c["n"], c["n"]), dtype=np.int64)
        assert Matrix(A.tolist()).determinant() == int(c["det"])


def test_builders_and_eigen_callers_match_reference_goldens():
    """SURVEY.md section 8f items 1 and 2 against outputs of the UNMODIFIED reference (tests/golden/eig_builders,
    oracle/gen_golden.py eig): under the same random.seed our gen_unimodular_matrix / build_diagonalizable /
    build_jordanized (device inverse of P) give the reference's matrices; find_eigenspace (device kernel) gives the
    reference's particular vector, generators and dimension for every eigenvalue; and the P, P^-1 and P^-1 A P that
    diagonalize() forms from them (device inverse) are the reference's."""
    import random
    import sympy
    from linalg_solver_b200 import Matrix, RandomMatrixBuilder
    from linalg_solver_b200.random_matrix import gen_unimodular_matrix
    g = golden_io.load("eig_builders")
    flat = lambda M: [pq(x) for row in M.items for x in row]
    n_diag = 0
    for c in g["cases"]:
        N = c["N"]
        random.seed(c["seed"])
        assert flat(gen_unimodular_matrix(N)) == c["unimodular"]
        random.seed(c["seed"])
        b = RandomMatrixBuilder.new().with_size(N, N)
        spec = [tuple(x) for x in c["spec"]]
        b = b.with_eigenvalues(spec) if c["kind"] == "diag" else b.with_jordan_blocks(spec)
        A = b.build()
        assert flat(A) == c["A"], (c["kind"], c["spec"])
        basis = []
        for sp_want in c["eigenspaces"]:
            sp = A.find_eigenspace(sympy.Integer(sp_want["eig"]))
            assert sp.dim() == sp_want["dim"] and [pq(x) for x in sp.vec] == sp_want["vec"]
            assert (flat(sp.generators) if sp.generators.cols else []) == sp_want["generators"]
            basis.extend(sp.basis())
        assert (len(basis) == N) == ("P" in c)
        if "P" in c:
            n_diag += 1
            P = Matrix([list(col) for col in zip(*basis)])
            P_inv = P.inverse()
            assert flat(P) == c["P"] and flat(P_inv) == c["P_inv"] and flat(P_inv * A * P) == c["D"]
            res = A.diagonalize()
            assert res.success and sorted(pq(res.D.items[i][i])[0] for i in range(N)) == \
                sorted(e for e, mlt in spec for _ in range(mlt))
        else:
            assert not A.diagonalize().success
    assert n_diag == 8


def test_row_reduce_trace_of_rational_matrices():
    """row_reduce(trace=True) on Fraction entries (reference linalg.py:534-630 accepts any exact entries): steps and
    frames against the unmodified reference (tests/golden/trace_rational); the device replays on the residues of
    numerators / common denominator (lsx_rref_trace_q)."""
    from linalg_solver_b200 import Matrix
    g = golden_io.load("trace_rational")
    seen = set()
    for c in g["cases"]:
        M = Matrix([[Fraction(p, q) for p, q in row] for row in c["A"]])
        R, piv, mats, steps = M.row_reduce(c["bar_col"], trace=True) if c["bar_col"] is not None else M.row_reduce(trace=True)
        assert [list(s) for s in steps] == c["steps"], (c["A"], c["bar_col"])
        assert [[pq(x) for row in f for x in row] for f in mats] == c["frames"]
        assert [pq(x) for row in R for x in row] == c["rref"] and [list(p) for p in piv] == c["pivots"]
        seen |= {s[0][0] for s in steps}
    assert seen == {"S", "N", "E"}


def test_trace_votes_out_a_bad_prime():
    """a prime that divides an intermediate value logs different steps; it is dropped and replaced instead
    of failing the call.  Tiny first primes (5 divides the pivot 5; 7 divides the denominator of a later entry)."""
    from linalg_solver_b200 import Matrix, default_engine
    eng = default_engine()
    A = [[5, 1, 2], [3, 7, 1], [2, 2, 9]]
    want = Matrix(A).row_reduce(trace=True)
    eng.debug_set_primes([5, 7, 11])
    try:
        got = Matrix(A).row_reduce(trace=True)
    finally:
        eng.debug_set_primes([])
    assert got[1] == want[1] and got[3] == want[3]
    assert [[pq(x) for row in f for x in row] for f in got[2]] == [[pq(x) for row in f for x in row] for f in want[2]]


def test_rank_beyond_the_batched_kernels():
    """rank() past the shared-memory tile (m > 254) through lsx_rank_large: DomainMatrix stand-ins (tests/golden/
    rank_large) and constructions whose rank is certain ([I; X] [I | Y] has rank exactly r), rows and columns
    shuffled; 1024 x 1024 included.  Also a small matrix with entries that need more than 32 primes."""
    import numpy as np
    from linalg_solver_b200 import Matrix, default_engine
    eng = default_engine()
    for c in golden_io.load("rank_large")["cases"]:
        rng = np.random.Generator(np.random.PCG64(c["seed"]))
        B = rng.integers(-3, 4, size=(c["m"], c["r"]), dtype=np.int64)
        C = rng.integers(-3, 4, size=(c["r"], c["n"]), dtype=np.int64)
        A = (B @ C).astype(np.int32)
        got, used = eng.rank_large(A)
        assert got == c["rank"], (c, got)
        assert used >= 1
    rng = np.random.Generator(np.random.PCG64(99))
    for m, n, r in [(300, 300, 300), (400, 333, 57), (1024, 1024, 1000), (1024, 1024, 1024), (255, 600, 1)]:
        B = np.vstack([np.eye(r, dtype=np.int64), rng.integers(-2, 3, size=(m - r, r))]) if m > r else np.eye(r, dtype=np.int64)[:m]
        C = np.hstack([np.eye(r, dtype=np.int64), rng.integers(-2, 3, size=(r, n - r))]) if n > r else np.eye(r, dtype=np.int64)[:, :n]
        A = (B @ C)[rng.permutation(m)][:, rng.permutation(n)].astype(np.int32)
        import torch
        got_h, _ = eng.rank_large(A)
        got_d, _ = eng.rank_large(torch.from_numpy(A).cuda())
        assert got_h == got_d == min(r, m, n), (m, n, r, got_h, got_d)
    assert eng.rank_large(np.zeros((300, 10), dtype=np.int32)) == (0, 0)
    # through the Matrix API: 260 x 260 (rows > 254) and a 6 x 6 whose entries need more primes than the batched limit
    A = rng.integers(-5, 6, size=(260, 260), dtype=np.int64)
    A[100] = A[3] - 2 * A[7]
    assert Matrix(A.tolist()).rank() == 259
    big = rng.integers(-2**30, 2**30, size=(40, 40), dtype=np.int64)
    big[5] = big[6]
    assert Matrix(big.tolist()).rank() == 39
