"""Pin the CPU oracle (oracle/ref_port.py) to outputs of the reference itself.

The fixtures under tests/golden/ were produced by oracle/gen_golden.py running the
unmodified reference (linalg.py:534-999) in the build container.  CPU only.
"""
from fractions import Fraction

import pytest

from oracle import golden_io, ref_port


def pq_list(grid):
    return [[x.numerator, x.denominator] for row in grid for x in row]


def affine_json(res, nvars, route):
    """Mirror gen_golden.affine_to_json for an oracle result."""
    if res is None:
        return {"status": "nosolution"}
    part, gens = res
    k = len(gens)
    if route == "default":
        order = ref_port.sympy_generator_order(k)
        gens = [gens[i] for i in order]
        gcols = k
    else:
        gcols = k if k else None
    flat_part = [[x.numerator, x.denominator] for x in part]
    flat_g = [[gens[c][r].numerator, gens[c][r].denominator] for r in range(nvars) for c in range(k)]
    return {"status": "ok", "gen_cols": gcols, "sha": golden_io.digest_pq(flat_part + flat_g),
            "particular": flat_part, "generators": flat_g}


def check_affine(case):
    n = len(case["A"][0])
    res = ref_port.find_preimage_of(case["A"], case["b"])
    for route in ("default", "logged"):
        want = case[route]
        got = affine_json(res, n, route)
        assert got["status"] == want["status"]
        if want["status"] != "ok":
            continue
        assert got["gen_cols"] == want["gen_cols"], route
        assert got["sha"] == want["sha"], route
        if "particular" in want:
            assert got["particular"] == want["particular"]
            assert got["generators"] == want["generators"]


def test_edge_row_reduce():
    g = golden_io.load("edge_small")
    assert len(g["rref_cases"]) > 400
    for c in g["rref_cases"]:
        R, piv = ref_port.row_reduce(c["A"], c["bar_col"])
        assert pq_list(R) == c["rref"], (c["A"], c["bar_col"])
        assert [list(p) for p in piv] == c["pivots"]
        assert ref_port.rank(c["A"]) == c["rank"]


def test_edge_systems():
    g = golden_io.load("edge_small")
    for c in g["system_cases"]:
        check_affine(c)


def test_edge_inverse_and_det():
    g = golden_io.load("edge_small")
    for c in g["inverse_cases"]:
        inv = ref_port.inverse(c["A"])
        got = None if inv is None else pq_list(inv)
        assert got == c["default"]
        assert got == c["logged"]
        d = ref_port.determinant(c["A"])
        assert [d.numerator, d.denominator] == c["det"]
        assert ref_port.bareiss_det(c["A"]) == c["det"][0]


def test_c1_all():
    g = golden_io.load("c1_4x4")
    assert len(g["cases"]) == 10000
    nsing = 0
    for c in g["cases"]:
        R, piv = ref_port.row_reduce(c["A"])
        assert pq_list(R) == c["rref"]
        assert [list(p) for p in piv] == c["pivots"]
        assert ref_port.rank(c["A"]) == c["rank"]
        d = ref_port.determinant(c["A"])
        assert [d.numerator, d.denominator] == c["det"]
        nsing += c["rank"] < 4
    assert nsing > 0


def test_c2_sample():
    g = golden_io.load("c2_8x8")
    cases = g["cases"]
    assert len(cases) == 2048 + 8
    for i, c in enumerate(cases):
        if i >= 320 and i < 2048:
            continue        # keep the CPU suite short; the GPU suite checks all of them
        inv = ref_port.inverse(c["A"])
        if "inverse_sha" in c:
            assert golden_io.digest_pq(pq_list(inv)) == c["inverse_sha"]
            if "inverse" in c and c["inverse"] is not None:
                assert pq_list(inv) == c["inverse"]
        else:
            assert inv is None
        if "rref_aug" in c:
            n = len(c["A"])
            aug = [list(c["A"][r]) + [1 if r == k else 0 for k in range(n)] for r in range(n)]
            R, piv = ref_port.row_reduce(aug, n)
            assert pq_list(R) == c["rref_aug"]
            assert [list(p) for p in piv] == c["pivots"]
            assert c["logged_inverse_equal"]
        if "det" in c:
            d = ref_port.determinant(c["A"])
            assert [d.numerator, d.denominator] == c["det"]


def test_c3_sample():
    g = golden_io.load("c3_16x17")
    assert len(g["cases"]) == 1024
    n_incons = 0
    for c in g["cases"][:160]:
        check_affine(c)
        n_incons += c["default"]["status"] != "ok"
    assert n_incons > 0


def test_c4_sample():
    g = golden_io.load("c4_64x64")
    c = g["inverse_cases"][0]
    inv = ref_port.inverse(c["A"])
    assert golden_io.digest_pq(pq_list(inv)) == c["inverse_sha"]
    assert c["rref_right_equals_inverse"]
    k = g["kernel_cases"][0]
    res = ref_port.kernel(k["A"])
    got = affine_json(res, 64, "default")
    assert got["gen_cols"] == k["gen_cols"]
    assert got["sha"] == k["sha"]


def test_c5_standins_bareiss():
    import numpy as np
    g = golden_io.load("c5_standins")
    for c in g["cases"]:
        if c["n"] > 128:
            continue
        rng = np.random.Generator(np.random.PCG64(c["seed"]))
        A = rng.integers(-5, 6, size=(c["n"], c["n"]), dtype=np.int64).tolist()
        assert ref_port.bareiss_det(A) == int(c["det"])


def test_det_mod_p_oracle_pinned_to_standins():
    """oracle/det_mod_p.py (numpy modular elimination, the config 5 checker) against the exact DomainMatrix
    determinants, for table primes and a tiny prime (zero pivots, swaps)."""
    import numpy as np
    from oracle.det_mod_p import det_mod_p
    from tests.device_model import prime_table
    g = golden_io.load("c5_standins")
    primes = prime_table(3) + [7, 65537]
    for c in g["cases"]:
        rng = np.random.Generator(np.random.PCG64(c["seed"]))
        A = rng.integers(-5, 6, size=(c["n"], c["n"]), dtype=np.int64)
        for p in primes:
            assert det_mod_p(A, p) == int(c["det"]) % p
    S = np.array([[1, 2, 3], [2, 4, 6], [1, 0, 1]])
    assert det_mod_p(S, primes[0]) == 0


def test_row_reduce_trace_pinned_to_reference_frames():
    """oracle row_reduce_trace (step labels, descriptions and every intermediate matrix) against the frames the
    unmodified reference produced (tests/golden/trace_small, generated by oracle/gen_golden.py trace)."""
    g = golden_io.load("trace_small")
    kinds = set()
    for c in g["cases"]:
        R, piv, frames, steps = ref_port.row_reduce_trace(c["A"], c["bar_col"])
        assert [list(s) for s in steps] == c["steps"], c["A"]
        assert [pq_list(f) for f in frames] == c["frames"]
        assert pq_list(R) == c["rref"] and [list(p) for p in piv] == c["pivots"]
        kinds |= {s[0][0] for s in steps}
    assert kinds == {"S", "N", "E"}


def test_row_reduce_trace_of_rational_matrices_pinned():
    """The oracle's step trace on fractional entries against the unmodified reference (tests/golden/trace_rational)."""
    from fractions import Fraction
    g = golden_io.load("trace_rational")
    for c in g["cases"]:
        items = [[Fraction(p, q) for p, q in row] for row in c["A"]]
        R, piv, frames, steps = ref_port.row_reduce_trace(items, c["bar_col"])
        assert [list(s) for s in steps] == c["steps"], c["A"]
        assert [pq_list(f) for f in frames] == c["frames"]
        assert pq_list(R) == c["rref"] and [list(p) for p in piv] == c["pivots"]
