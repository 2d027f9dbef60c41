#!/usr/bin/env python
"""Generate tests/golden/*.json.gz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs in the build container, where /root/reference
exists; the GPU box never runs this script (it only reads the committed
fixtures).  The reference package is imported from /root/reference with
``oracle/standin/linalg_helper.py`` standing in for its Rust module.  Inputs
come from the reference's own ``RandomMatrixBuilder`` under fixed seeds and are
rationalised with ``sympy.Rational`` exactly as reference main.py:20-31 does,
because raw ints make ``row_reduce`` fall into floats (linalg.py:574).

Usage:  python oracle/gen_golden.py [c1 c2 c3 c4 edge c5 trace eig traceq ranklarge]   (default: all)
"""
import os
import random
import sys
import time
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "standin"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import sympy  # noqa: E402
from linalg_solver.linalg import Matrix  # noqa: E402
from linalg_solver.log import global_logger, nest_logger  # noqa: E402
from linalg_solver.random_matrix import RandomMatrixBuilder  # noqa: E402

import golden_io  # noqa: E402

global_logger._auto_print = False
NPROC = min(8, os.cpu_count() or 1)


def rat(items):
    return [[sympy.Rational(x) for x in row] for row in items]


def pq(x):
    x = sympy.Rational(x)
    return [int(x.p), int(x.q)]


def pq_grid(rows):
    return [pq(x) for row in rows for x in row]


def quiet(f):
    with nest_logger():
        return f()


# ----------------------------------------------------------------------------- C1
def c1_case(items):
    M = Matrix(rat(items))
    det = quiet(lambda: M.determinant(use_optimal=False))
    rk = M.rank()
    R, piv, _, _ = M.row_reduce()
    return {"A": items, "det": pq(det), "rank": int(rk), "rref": pq_grid(R),
            "pivots": [list(p) for p in piv]}


def gen_c1():
    random.seed(20260001)
    mats = [RandomMatrixBuilder.new().with_size(4, 4).build().items for _ in range(10000)]
    with Pool(NPROC) as pool:
        cases = pool.map(c1_case, mats, chunksize=64)
    golden_io.save("c1_4x4", {
        "about": "reference Matrix.determinant(use_optimal=False), rank(), row_reduce() (default bar_col -> 3) on 10k RandomMatrixBuilder 4x4, random.seed(20260001)",
        "cases": cases})


# ----------------------------------------------------------------------------- C2
def c2_case(arg):
    idx, items = arg
    n = len(items)
    M = Matrix(rat(items))
    inv = M.inverse()
    out = {"A": items}
    if isinstance(inv, Matrix.NoSolution):
        out["inverse"] = None
    else:
        flat = pq_grid(inv.items)
        out["inverse_sha"] = golden_io.digest_pq(flat)
        if idx < 64:
            out["inverse"] = flat
    if idx < 256:
        aug = [list(M.items[i]) + [sympy.Integer(1 if i == j else 0) for j in range(n)] for i in range(n)]
        R, piv, _, _ = Matrix(aug).row_reduce(bar_col=n)
        out["rref_aug"] = pq_grid(R)
        out["pivots"] = [list(p) for p in piv]
        logged = quiet(lambda: M.inverse(log_result=True))
        out["logged_inverse_equal"] = (
            isinstance(logged, Matrix.NoSolution) == isinstance(inv, Matrix.NoSolution)
            and (isinstance(inv, Matrix.NoSolution) or pq_grid(logged.items) == pq_grid(inv.items)))
    if idx < 16:
        out["det"] = pq(quiet(lambda: M.determinant(use_optimal=False)))
    return out


def gen_c2():
    random.seed(20260002)
    mats = [RandomMatrixBuilder.new().with_size(8, 8).build().items for _ in range(2048)]
    # a few deliberately singular 8x8 (duplicate / zero rows) appended after the seeded ones
    extra = []
    for k in range(8):
        A = [row[:] for row in mats[k]]
        if k % 2 == 0:
            A[5] = A[2][:]
        else:
            A[k] = [0] * 8
        extra.append(A)
    allm = mats + extra
    with Pool(NPROC) as pool:
        cases = pool.map(c2_case, [(i if i < 2048 else 0, m) for i, m in enumerate(allm)], chunksize=8)
    golden_io.save("c2_8x8", {
        "about": "reference inverse() (sympy route) on 2048 builder 8x8 (random.seed(20260002)) + 8 planted singular; row_reduce([A|I], bar_col=8) and logged inverse on the first 256; legacy determinant on the first 16",
        "cases": cases})


# ----------------------------------------------------------------------------- C3
def affine_to_json(res, full):
    if isinstance(res, Matrix.NoSolution):
        return {"status": "nosolution"}
    part = [pq(x) for x in res.vec]
    gm = res.generators
    if gm is None:
        gens, gcols = [], None
    else:
        gcols = gm.cols
        gens = pq_grid(gm.items)
    out = {"status": "ok", "gen_cols": gcols,
           "sha": golden_io.digest_pq(part + gens)}
    if full:
        out["particular"] = part
        out["generators"] = gens
    return out


def c3_build(seed):
    random.seed(seed)
    A = RandomMatrixBuilder.new().with_size(16, 16).with_rank(10).build().items
    if seed % 2 == 0:
        x0 = [random.randint(-5, 5) for _ in range(16)]
        b = [sum(A[i][j] * x0[j] for j in range(16)) for i in range(16)]
    else:
        b = [random.randint(-5, 5) for _ in range(16)]
    return A, b


def c3_case(arg):
    idx, seed = arg
    A, b = c3_build(seed)
    M = Matrix(rat(A))
    bv = [sympy.Rational(x) for x in b]
    dflt = M.find_preimage_of(list(bv))
    logged = quiet(lambda: Matrix(rat(A)).find_preimage_of(list(bv), log_result=True))
    return {"A": A, "b": b, "default": affine_to_json(dflt, idx < 64),
            "logged": affine_to_json(logged, idx < 64)}


def gen_c3():
    seeds = [202600030000 + i for i in range(1024)]
    with Pool(NPROC) as pool:
        cases = pool.map(c3_case, list(enumerate(seeds)), chunksize=4)
    golden_io.save("c3_16x17", {
        "about": "reference find_preimage_of (default sympy route and logged row_reduce route) on 1024 systems; A = RandomMatrixBuilder 16x16 with_rank(10) under random.seed(202600030000+i); even i: b = A*x0, odd i: b uniform [-5,5]",
        "cases": cases})


# ----------------------------------------------------------------------------- C4
def c4_inv_case(arg):
    idx, items = arg
    M = Matrix(rat(items))
    inv = M.inverse()
    out = {"A": items}
    if isinstance(inv, Matrix.NoSolution):
        out["inverse"] = None
        return out
    flat = pq_grid(inv.items)
    out["inverse_sha"] = golden_io.digest_pq(flat)
    out["inverse_row0"] = flat[:64]
    if idx < 2:
        # the row_reduce route itself, on Fractions (sympy.Rational would take minutes)
        from fractions import Fraction
        n = len(items)
        aug = [[Fraction(x) for x in items[i]] + [Fraction(1 if i == j else 0) for j in range(n)] for i in range(n)]
        R, piv, _, _ = Matrix(aug).row_reduce(bar_col=n)
        rflat = [[x.numerator, x.denominator] for row in R for x in row]
        out["rref_aug_sha"] = golden_io.digest_pq(rflat)
        out["pivots"] = [list(p) for p in piv]
        out["rref_right_equals_inverse"] = ([v for i in range(n) for v in rflat[i * 2 * n + n:(i + 1) * 2 * n]] == flat)
    return out


def c4_ker_case(arg):
    idx, items = arg
    M = Matrix(rat(items))
    res = M.kernel()
    out = {"A": items}
    out.update(affine_to_json(res, False))
    gm = res.generators
    out["gen_col0"] = [pq(x) for x in gm.get_col(0)] if gm is not None and gm.cols else []
    return out


def gen_c4():
    random.seed(20260004)
    inv_mats = [RandomMatrixBuilder.new().with_size(64, 64).build().items for _ in range(64)]
    ker_mats = []
    for _ in range(16):
        B = RandomMatrixBuilder.new().with_size(64, 48).build().items
        C = RandomMatrixBuilder.new().with_size(48, 64).build().items
        ker_mats.append([[sum(B[i][k] * C[k][j] for k in range(48)) for j in range(64)] for i in range(64)])
    with Pool(NPROC) as pool:
        inv_cases = pool.map(c4_inv_case, list(enumerate(inv_mats)), chunksize=1)
        ker_cases = pool.map(c4_ker_case, list(enumerate(ker_mats)), chunksize=1)
    golden_io.save("c4_64x64", {
        "about": "reference inverse() on 64 builder 64x64 (random.seed(20260004)); row_reduce([A|I],64) on Fractions for the first 2; kernel() (sympy route) on 16 products B(64x48)*C(48x64) of builder matrices (no rank() rejection: infeasible at this size)",
        "inverse_cases": inv_cases, "kernel_cases": ker_cases})


# ----------------------------------------------------------------------------- edge
def edge_inputs():
    rnd = random.Random(20260099)
    cases = []
    shapes = [(1, 1), (1, 2), (2, 1), (1, 3), (2, 2), (2, 3), (3, 2), (3, 3), (3, 4), (4, 3), (3, 5), (5, 3), (4, 4),
              (4, 5), (5, 4), (5, 5), (4, 6), (6, 4), (6, 6), (5, 7), (7, 5), (2, 6), (6, 2), (8, 9), (9, 8), (3, 8)]
    for (m, n) in shapes:
        bars = [None, 0] + list(range(1, n + 1))
        for bar in bars:
            for style in range(3):
                if style == 0:      # dense small
                    A = [[rnd.randint(-3, 3) for _ in range(n)] for _ in range(m)]
                elif style == 1:    # sparse: many zeros -> swaps and skipped columns
                    A = [[rnd.choice([0, 0, 0, 1, -1, 2]) for _ in range(n)] for _ in range(m)]
                else:               # low rank: every row a combination of two rows
                    u = [rnd.randint(-2, 2) for _ in range(n)]
                    v = [rnd.randint(-2, 2) for _ in range(n)]
                    A = [[rnd.randint(-2, 2) * u[j] + rnd.randint(-1, 1) * v[j] for j in range(n)] for _ in range(m)]
                cases.append((A, bar))
    # all-zero, identity-like, permutation matrices
    cases.append(([[0, 0, 0], [0, 0, 0]], None))
    cases.append(([[0, 0, 0], [0, 0, 0]], 3))
    cases.append(([[0, 1, 0], [0, 0, 1], [1, 0, 0]], 3))
    cases.append(([[0, 0, 1], [0, 1, 0], [1, 0, 0]], 3))
    cases.append(([[1, 1, 0], [1, 0, 1]], 1))
    cases.append(([[1, 0, 1], [1, 1, 0]], 1))
    return cases


def edge_rref_case(arg):
    A, bar = arg
    R, piv, _, _ = Matrix(rat(A)).row_reduce(bar_col=bar) if bar is not None else Matrix(rat(A)).row_reduce()
    return {"A": A, "bar_col": bar, "rref": pq_grid(R), "pivots": [list(p) for p in piv],
            "rank": int(Matrix(rat(A)).rank())}


def edge_solver_inputs():
    rnd = random.Random(20260098)
    sys_cases, inv_cases = [], []
    for (m, n) in [(2, 2), (3, 3), (3, 4), (4, 3), (4, 4), (5, 5), (4, 6), (6, 4), (6, 6), (2, 5), (5, 2), (1, 1), (1, 4)]:
        for style in range(6):
            if style < 2:
                A = [[rnd.randint(-4, 4) for _ in range(n)] for _ in range(m)]
            elif style < 4:
                u = [rnd.randint(-2, 2) for _ in range(n)]
                A = [[rnd.randint(-2, 2) * x for x in u] for _ in range(m)]
                A[0] = [rnd.randint(-1, 1) for _ in range(n)]
            else:
                A = [[rnd.choice([0, 0, 1, -1]) for _ in range(n)] for _ in range(m)]
            if style % 2 == 0:
                x0 = [rnd.randint(-3, 3) for _ in range(n)]
                b = [sum(A[i][j] * x0[j] for j in range(n)) for i in range(m)]
            else:
                b = [rnd.randint(-3, 3) for _ in range(m)]
            sys_cases.append((A, b))
    sys_cases.append(([[0, 0], [0, 0]], [0, 0]))
    sys_cases.append(([[0, 0], [0, 0]], [0, 1]))
    # > 10 free variables: generator order follows sorted(str(tau_k)) on the default route
    u = [1, 2, -1, 3, 0, 1, -2, 1, 1, 0, 2, -1, 1, 3]
    v = [0, 1, 1, -1, 2, 0, 1, 3, -1, 1, 0, 2, 1, 1]
    A14 = [[(i % 3 - 1) * u[j] + ((i * 7) % 5 - 2) * v[j] for j in range(14)] for i in range(14)]
    sys_cases.append((A14, [0] * 14))
    sys_cases.append((A14, [A14[i][0] - 2 * A14[i][5] for i in range(14)]))
    for n in [1, 2, 3, 4, 5, 6]:
        for style in range(4):
            if style < 2:
                A = [[rnd.randint(-4, 4) for _ in range(n)] for _ in range(n)]
            elif style == 2:
                A = [[rnd.choice([0, 0, 1, -1, 2]) for _ in range(n)] for _ in range(n)]
            else:
                A = [[rnd.randint(-3, 3) for _ in range(n)] for _ in range(n)]
                A[n - 1] = A[0][:]
            inv_cases.append(A)
    return sys_cases, inv_cases


def edge_sys_case(arg):
    A, b = arg
    bv = [sympy.Rational(x) for x in b]
    dflt = Matrix(rat(A)).find_preimage_of(list(bv))
    logged = quiet(lambda: Matrix(rat(A)).find_preimage_of(list(bv), log_result=True))
    return {"A": A, "b": b, "default": affine_to_json(dflt, True), "logged": affine_to_json(logged, True)}


def edge_inv_case(A):
    M = Matrix(rat(A))
    dflt = M.inverse()
    logged = quiet(lambda: Matrix(rat(A)).inverse(log_result=True))
    n = len(A)
    det = quiet(lambda: Matrix(rat(A)).determinant(use_optimal=False))

    def enc(r):
        return None if isinstance(r, Matrix.NoSolution) else pq_grid(r.items)
    return {"A": A, "default": enc(dflt), "logged": enc(logged), "det": pq(det), "n": n}


def gen_edge():
    rcases = edge_inputs()
    scases, icases = edge_solver_inputs()
    with Pool(NPROC) as pool:
        r = pool.map(edge_rref_case, rcases, chunksize=8)
        s = pool.map(edge_sys_case, scases, chunksize=2)
        i = pool.map(edge_inv_case, icases, chunksize=2)
    golden_io.save("edge_small", {
        "about": "reference row_reduce over shapes 1x1..9x8 with every bar_col (None, 0, 1..n), dense/sparse/low-rank integer inputs; find_preimage_of (both routes) incl. inconsistent, unique, zero and >10-free-variable systems; inverse (both routes) + legacy determinant for n<=6",
        "rref_cases": r, "system_cases": s, "inverse_cases": i})


# ----------------------------------------------------------------------------- step traces
def trace_case(arg):
    """The reference's own row_reduce with its frame renderer wrapped: every call of
    make_latex_augmented_matrix(A, bar_col) inside linalg.py:534-630 is one recorded frame, so the exact rational
    intermediate matrices come from the unmodified reference (the LaTeX it returns is not kept)."""
    import copy
    import linalg_solver.linalg as L
    items, bar = arg
    frames = []
    orig = L.make_latex_augmented_matrix

    def spy(A, bar_col=None, **kw):
        frames.append(pq_grid(copy.deepcopy(A)))
        return orig(A, bar_col=bar_col, **kw)

    L.make_latex_augmented_matrix = spy
    try:
        M = Matrix(rat(items))
        R, piv, mats, steps = M.row_reduce(bar) if bar is not None else M.row_reduce()
    finally:
        L.make_latex_augmented_matrix = orig
    assert len(frames) == len(mats) == len(steps) + 1
    return {"A": items, "bar_col": bar, "steps": [list(x) for x in steps], "frames": frames,
            "rref": pq_grid(R), "pivots": [list(x) for x in piv]}


def gen_trace():
    rnd = random.Random(20260033)
    cases = []
    for m, n in [(1, 2), (2, 2), (2, 3), (3, 3), (3, 4), (4, 4), (4, 5), (3, 6), (5, 4), (5, 6), (6, 6), (6, 7), (8, 9)]:
        for rep in range(12):
            if rep % 4 == 3:                                      # low rank
                rk = max(1, min(m, n) - 1 - rep % 2)
                Bm = [[rnd.randint(-3, 3) for _ in range(rk)] for _ in range(m)]
                Cm = [[rnd.randint(-3, 3) for _ in range(n)] for _ in range(rk)]
                items = [[sum(Bm[i][k] * Cm[k][j] for k in range(rk)) for j in range(n)] for i in range(m)]
            else:
                items = [[rnd.randint(-5, 5) for _ in range(n)] for _ in range(m)]
            if rep % 4 == 1:
                items[0][0] = 0                                   # forces a swap (or a skipped column)
                if m > 1 and rep % 8 == 5:
                    items[1][0] = 0
            if rep % 4 == 2 and m > 1:
                items[0][0] = 1                                   # pivot already one: no N step
            bar = [None, n, max(1, n - 1), min(n, m)][rep % 4]
            cases.append((items, bar))
    with Pool(NPROC) as pool:
        out = pool.map(trace_case, cases, chunksize=4)
    golden_io.save("trace_small", {
        "about": "reference Matrix.row_reduce (linalg.py:534-630) with its frame renderer wrapped: step labels/descriptions "
                 "(intermediate_steps) and the exact rational intermediate matrices behind intermediate_matrices",
        "cases": out})


# ----------------------------------------------------------------------------- eigen-callers and builders
EIG_SPECS = [
    # (kind, N, spec): the reference builder under random.seed(20260060 + index)
    ("diag", 3, [(2, 1), (-1, 2)]), ("diag", 4, [(1, 2), (3, 2)]), ("diag", 4, [(0, 1), (2, 1), (-3, 2)]),
    ("diag", 5, [(1, 1), (2, 1), (3, 1), (-1, 2)]), ("diag", 5, [(4, 5)]), ("diag", 6, [(0, 2), (1, 2), (-2, 2)]),
    ("diag", 6, [(5, 1), (-5, 1), (2, 4)]), ("diag", 8, [(1, 3), (-1, 3), (2, 2)]),
    ("jordan", 3, [(2, 2), (1, 1)]), ("jordan", 4, [(1, 2), (1, 2)]), ("jordan", 4, [(0, 3), (2, 1)]),
    ("jordan", 5, [(3, 2), (3, 1), (-1, 2)]), ("jordan", 6, [(1, 3), (2, 3)]), ("jordan", 6, [(-2, 1), (-2, 2), (4, 3)]),
    ("jordan", 8, [(0, 4), (1, 2), (1, 2)]),
]


def eig_case(arg):
    idx, (kind, N, spec) = arg
    seed = 20260060 + idx
    random.seed(seed)
    P = quiet(lambda: __import__("linalg_solver.random_matrix", fromlist=["x"]).gen_unimodular_matrix(N))
    random.seed(seed)
    b = RandomMatrixBuilder.new().with_size(N, N)
    b = b.with_eigenvalues(spec) if kind == "diag" else b.with_jordan_blocks(spec)
    A = quiet(lambda: b.build())                                   # P^-1 D P / P^-1 J P through the reference's inverse
    eigs = sorted({e for e, _ in spec})
    spaces, basis = [], []
    for e in eigs:
        sp = quiet(lambda: A.find_eigenspace(sympy.Integer(e)))      # kernel -> find_preimage_of default route
        gens = sp.generators
        spaces.append({"eig": e, "vec": [pq(x) for x in sp.vec], "dim": int(sp.dim()),
                       "generators": pq_grid(gens.items) if gens.cols else []})
        basis.extend(sp.basis())
    out = {"kind": kind, "N": N, "spec": [list(x) for x in spec], "seed": seed, "unimodular": pq_grid(P.items),
           "A": pq_grid(A.items), "eigenspaces": spaces}
    if len(basis) == N:                                            # what diagonalize() does next (linalg.py:852-858)
        Pm = Matrix([list(col) for col in zip(*basis)])
        P_inv = quiet(lambda: Pm.inverse())
        out["P"] = pq_grid(Pm.items)
        out["P_inv"] = pq_grid(P_inv.items)
        out["D"] = pq_grid(quiet(lambda: P_inv * A * Pm).items)
    return out


def gen_eig():
    with Pool(NPROC) as pool:
        cases = pool.map(eig_case, list(enumerate(EIG_SPECS)), chunksize=1)
    golden_io.save("eig_builders", {
        "about": "reference gen_unimodular_matrix, RandomMatrixBuilder.build_diagonalizable / build_jordanized "
                 "(random_matrix.py:131-167, 233-267) under random.seed(20260060 + index); find_eigenspace "
                 "(linalg.py:758-770) for every eigenvalue in ascending order; P from the eigenvectors as columns, "
                 "P.inverse() and P^-1 A P as diagonalize() forms them (linalg.py:852-858)",
        "cases": cases})


# ----------------------------------------------------------------------------- step traces of RATIONAL matrices
def gen_trace_q():
    from fractions import Fraction
    rnd = random.Random(20260034)
    cases = []
    for m, n in [(2, 2), (2, 3), (3, 3), (3, 4), (4, 4), (4, 5), (5, 6)]:
        for rep in range(6):
            items = [[Fraction(rnd.randint(-6, 6), rnd.choice([1, 2, 3, 4, 6])) for _ in range(n)] for _ in range(m)]
            if rep % 3 == 1:
                items[0][0] = Fraction(0)
            if rep % 3 == 2:
                items[0][0] = Fraction(1)                          # pivot already one: no N step
            bar = [None, n, max(1, n - 1)][rep % 3]
            cases.append(([[[x.numerator, x.denominator] for x in row] for row in items], bar))
    with Pool(NPROC) as pool:
        out = pool.map(trace_q_case, cases, chunksize=4)
    golden_io.save("trace_rational", {
        "about": "as trace_small, for matrices with fractional entries (row_reduce accepts any exact entries, linalg.py:534-630)",
        "cases": out})


def trace_q_case(arg):
    items, bar = arg
    res = trace_case(([[sympy.Rational(p, q) for p, q in row] for row in items], bar))
    res["A"] = items
    return res


# ----------------------------------------------------------------------------- C5 stand-ins
def gen_c5():
    """No reference route can compute these; third-party cross-oracle only."""
    import numpy as np
    from sympy.polys.matrices import DomainMatrix
    from sympy import ZZ
    out = []
    for n, seed in [(64, 2026000564), (128, 2026000528), (256, 2026000556), (384, 2026000538), (512, 2026000551)]:   # 384/512: 1.5 / 6 minutes
        rng = np.random.Generator(np.random.PCG64(seed))
        A = rng.integers(-5, 6, size=(n, n), dtype=np.int64)
        dm = DomainMatrix([[ZZ(int(x)) for x in row] for row in A.tolist()], (n, n), ZZ)
        out.append({"n": n, "seed": seed, "det": str(int(dm.det()))})
    golden_io.save("c5_standins", {
        "about": "NOT from the reference (its determinant is infeasible for dense n >= 9): sympy DomainMatrix(ZZ).det() of numpy PCG64(seed).integers(-5,6,(n,n)) as an independent exact cross-check",
        "cases": out})


# ----------------------------------------------------------------------------- large-n rank stand-ins
def gen_ranklarge():
    """Not from the reference (its rank() is sympy.Matrix.rank, minutes to hours at these sizes):
    sympy DomainMatrix(ZZ).rank() of products B (m x r) C (r x n) of numpy PCG64 matrices, as an exact cross-check of
    the sizes beyond the batched kernels (m > 254)."""
    import numpy as np
    from sympy.polys.matrices import DomainMatrix
    from sympy import ZZ
    out = []
    for m, n, r, seed in [(260, 270, 200, 2026000701), (300, 280, 280, 2026000702), (512, 512, 100, 2026000703)]:
        rng = np.random.Generator(np.random.PCG64(seed))
        B = rng.integers(-3, 4, size=(m, r), dtype=np.int64)
        C = rng.integers(-3, 4, size=(r, n), dtype=np.int64)
        A = B @ C
        dm = DomainMatrix([[ZZ(int(x)) for x in row] for row in A.tolist()], (m, n), ZZ)
        out.append({"m": m, "n": n, "r": r, "seed": seed, "rank": int(dm.rank())})
    golden_io.save("rank_large", {
        "about": "NOT from the reference: sympy DomainMatrix(ZZ).rank() of (PCG64(seed).integers(-3,4,(m,r)) @ "
                 "PCG64-continued .integers(-3,4,(r,n))) as an independent exact cross-check",
        "cases": out})


ALL = {"c1": gen_c1, "c2": gen_c2, "c3": gen_c3, "c4": gen_c4, "edge": gen_edge, "c5": gen_c5, "trace": gen_trace,
       "eig": gen_eig, "traceq": gen_trace_q, "ranklarge": gen_ranklarge}

if __name__ == "__main__":
    which = sys.argv[1:] or list(ALL)
    for w in which:
        t0 = time.time()
        ALL[w]()
        print("%s done in %.1fs" % (w, time.time() - t0), flush=True)
