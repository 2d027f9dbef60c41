"""Parity of the CUDA path (through the C-ABI, host buffers) with the reference-generated golden
fixtures and the CPU oracle.  Bit-exact: every comparison is on exact integers / (p, q) pairs."""
from fractions import Fraction

import numpy as np
import pytest

from oracle import golden_io, ref_port
from tests.helpers import limbs_to_ints, np_batch, pq_grid_from_num, pq_of_fracs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from linalg_solver_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def run_rref(eng, mats, bar, **kw):
    res = eng.rref_batch(np_batch(mats), bar, **kw)
    assert not np.any(res.status & ~32), res.status[np.nonzero(res.status & ~32)]
    num, den = limbs_to_ints(res.num), limbs_to_ints(res.den)
    return res, num, den


def check_rref_against_oracle(eng, mats, bar):
    res, num, den = run_rref(eng, mats, bar)
    for i, A in enumerate(mats):
        R, piv = ref_port.row_reduce(A, bar)
        assert int(res.rank[i]) == len(piv)
        assert [int(c) for c in res.pivot_col[i][:len(piv)]] == [c for _, c in piv]
        assert all(int(c) == -1 for c in res.pivot_col[i][len(piv):])
        assert den[i] != 0
        assert pq_grid_from_num(num[i], den[i]) == pq_of_fracs(R), (A, bar)


# ------------------------------------------------------------------ edge cases (reference golden)
def test_edge_row_reduce_every_shape_and_bar(eng):
    g = golden_io.load("edge_small")
    groups = {}
    for c in g["rref_cases"]:
        A = c["A"]
        bar = c["bar_col"] or len(A[0]) - 1
        if bar <= 0:
            continue                                     # no pivot column: the wrapper returns the input
        groups.setdefault((len(A), len(A[0]), bar), []).append(c)
    assert len(groups) > 40
    for (m, n, bar), cases in groups.items():
        res, num, den = run_rref(eng, [c["A"] for c in cases], bar)
        for i, c in enumerate(cases):
            assert pq_grid_from_num(num[i], den[i]) == c["rref"], (c["A"], bar)
            rk = int(res.rank[i])
            assert [[k, int(res.pivot_col[i][k])] for k in range(rk)] == c["pivots"]


def test_edge_rank(eng):
    g = golden_io.load("edge_small")
    groups = {}
    for c in g["rref_cases"]:
        groups.setdefault((len(c["A"]), len(c["A"][0])), []).append(c)
    for cases in groups.values():
        res = eng.rank_batch(np_batch([c["A"] for c in cases]))
        assert [int(r) for r in res.rank] == [c["rank"] for c in cases]


def test_edge_inverse_and_det(eng):
    g = golden_io.load("edge_small")
    groups = {}
    for c in g["inverse_cases"]:
        groups.setdefault(c["n"], []).append(c)
    for n, cases in groups.items():
        A = np_batch([c["A"] for c in cases])
        res = eng.inverse_batch(A)
        adj, det = limbs_to_ints(res.adj), limbs_to_ints(res.det)
        dres = eng.det_batch(A)
        dets = limbs_to_ints(dres.det)
        for i, c in enumerate(cases):
            assert [dets[i], 1] == c["det"]
            assert det[i] == c["det"][0]
            if c["default"] is None:
                assert int(res.status[i]) & 1
                assert det[i] == 0 and all(x == 0 for r in adj[i] for x in r)
            else:
                assert int(res.status[i]) & 1 == 0
                assert pq_grid_from_num(adj[i], det[i]) == c["default"] == c["logged"]


def affine_pq(eng_res, i, n, route):
    """Mirror of gen_golden.affine_to_json for one device result."""
    st = int(eng_res.status[i])
    if st & 2:
        return {"status": "nosolution"}
    assert st & ~32 == 0
    den = limbs_to_ints(eng_res.den[i])
    part = limbs_to_ints(eng_res.particular[i])
    k = n - int(eng_res.rank[i])
    gens = limbs_to_ints(eng_res.generators[i]) if k else None
    order = ref_port.sympy_generator_order(k) if route == "default" else list(range(k))
    from linalg_solver_b200.convert import reduce_pq
    flat_part = [list(reduce_pq(x, den)) for x in part]
    flat_g = [list(reduce_pq(gens[r][c], den)) for r in range(n) for c in order]
    gcols = k if route == "default" else (k if k else None)
    return {"status": "ok", "gen_cols": gcols, "sha": golden_io.digest_pq(flat_part + flat_g),
            "particular": flat_part, "generators": flat_g}


def check_systems(eng, cases, **kw):
    groups = {}
    for c in cases:
        groups.setdefault((len(c["A"]), len(c["A"][0])), []).append(c)
    n_bad = 0
    for (m, n), cs in groups.items():
        res = eng.solve_batch(np_batch([c["A"] for c in cs]), np_batch([c["b"] for c in cs]), **kw)
        for i, c in enumerate(cs):
            for route in ("default", "logged"):
                want, got = c[route], affine_pq(res, i, n, route)
                assert got["status"] == want["status"], (c["A"], c["b"])
                if want["status"] != "ok":
                    n_bad += route == "default"
                    continue
                assert got["gen_cols"] == want["gen_cols"]
                assert got["sha"] == want["sha"]
                if "particular" in want:
                    assert got["particular"] == want["particular"]
                    assert got["generators"] == want["generators"]
    return n_bad


def test_edge_systems(eng):
    g = golden_io.load("edge_small")
    assert check_systems(eng, g["system_cases"]) > 0


# ------------------------------------------------------------------ config 1: 10k 4x4
def test_c1_all_10k(eng):
    g = golden_io.load("c1_4x4")
    cases = g["cases"]
    A = np_batch([c["A"] for c in cases])
    d = eng.det_batch(A, a_abs_max=5)
    dets = limbs_to_ints(d.det)
    rk = eng.rank_batch(A, a_abs_max=5)
    res, num, den = run_rref(eng, [c["A"] for c in cases], 3, a_abs_max=5, b_abs_max=5)
    nsing = 0
    for i, c in enumerate(cases):
        assert [dets[i], 1] == c["det"]
        assert int(rk.rank[i]) == c["rank"] == int(d.rank[i])
        assert pq_grid_from_num(num[i], den[i]) == c["rref"]
        assert [[k, int(res.pivot_col[i][k])] for k in range(int(res.rank[i]))] == c["pivots"]
        nsing += c["rank"] < 4
    assert nsing > 0


# ------------------------------------------------------------------ config 2: 8x8 det + inverse
def test_c2_golden(eng):
    g = golden_io.load("c2_8x8")
    cases = g["cases"]
    A = np_batch([c["A"] for c in cases])
    res = eng.inverse_batch(A, a_abs_max=5)
    adj, det = limbs_to_ints(res.adj), limbs_to_ints(res.det)
    n_sing = 0
    for i, c in enumerate(cases):
        if "inverse_sha" in c:
            assert int(res.status[i]) == 0
            pq = pq_grid_from_num(adj[i], det[i])
            assert golden_io.digest_pq(pq) == c["inverse_sha"]
            if c.get("inverse") is not None:
                assert pq == c["inverse"]
        else:
            assert int(res.status[i]) & 1
            n_sing += 1
        if "det" in c:
            assert [det[i], 1] == c["det"]
    assert n_sing >= 8
    # raw row_reduce([A|I], bar_col=8) including the planted singular ones (right block depends on
    # the reference's row choice)
    sub = [c for c in cases if "rref_aug" in c]
    aug = [[list(c["A"][r]) + [1 if r == k else 0 for k in range(8)] for r in range(8)] for c in sub]
    r2, num, den = run_rref(eng, aug, 8, a_abs_max=5, b_abs_max=1)
    for i, c in enumerate(sub):
        assert pq_grid_from_num(num[i], den[i]) == c["rref_aug"]
        assert [[k, int(r2.pivot_col[i][k])] for k in range(int(r2.rank[i]))] == c["pivots"]


def test_c2_full_size_properties(eng):
    """2^20 matrices: A * adj == det * I exactly (int64), singular <=> det == 0."""
    rng = np.random.Generator(np.random.PCG64(20260002))
    B = 1 << 20
    A = rng.integers(-5, 6, size=(B, 8, 8), dtype=np.int32)
    res = eng.inverse_batch(A, a_abs_max=5)
    assert res.plan.limbs == 1
    adj = res.adj.view(np.int32).reshape(B, 8, 8).astype(np.int64)
    det = res.det.view(np.int32).reshape(B).astype(np.int64)
    prod = np.einsum("bij,bjk->bik", A.astype(np.int64), adj)
    want = det[:, None, None] * np.eye(8, dtype=np.int64)[None]
    sing = (res.status & 1) != 0
    assert np.array_equal(prod[~sing], want[~sing])
    assert np.all(det[sing] == 0) and np.all(det[~sing] != 0)
    assert not np.any(res.status & ~1)
    # determinant cross-check on a sample against the oracle
    for i in range(0, B, B // 64):
        assert det[i] == ref_port.bareiss_det(A[i].tolist())
    assert sing.sum() > 0 or True


def _limbs_mod(t, q):
    """Signed multi-limb integers (torch int32 words, two's complement, last dim = limbs) reduced modulo a prime
    q < 2^20, on the tensor's device, as float64-exact integers."""
    import torch
    w = t.to(torch.int64) & 0xffffffff
    L = w.shape[-1]
    acc = torch.zeros(w.shape[:-1], dtype=torch.int64, device=w.device)
    for l in range(L - 1, -1, -1):
        acc = (acc * ((1 << 32) % q) + w[..., l] % q) % q
    neg = w[..., L - 1] >= (1 << 31)
    acc = torch.where(neg, (acc - pow(2, 32 * L, q)) % q, acc)
    return acc.to(torch.float64)


def _assert_canonical_solution_form(res, nvars, ok=None):
    """Canonical form of the reference's solution sets (linalg.py:961-983), on the EXACT device words and for every
    system of the batch: the particular solution is zero on the free columns; generator t is `den` at the t-th free
    column (ascending order), zero at the other free columns, and unused generator columns are zero."""
    import torch
    B = res.rank.shape[0]
    piv = res.pivot_col.to(torch.int64)                          # [B, slots], -1 padded
    is_piv = torch.zeros((B, nvars + 1), dtype=torch.bool, device=piv.device)
    is_piv.scatter_(1, torch.where(piv >= 0, piv, torch.full_like(piv, nvars)), True)
    free = ~is_piv[:, :nvars]                                    # [B, nvars]
    if ok is None:
        ok = torch.ones(B, dtype=torch.bool, device=piv.device)
    part_zero = (res.particular == 0).all(dim=-1)                # [B, nvars]
    assert bool((part_zero | ~free)[ok].all())
    G = res.generators.shape[2]
    order = torch.cumsum(free.to(torch.int64), dim=1) - 1        # index of a free column among the free columns
    gens = res.generators                                        # [B, nvars, G, L]
    den = res.den[:, None, None, :].expand(-1, nvars, G, -1)
    t_idx = torch.arange(G, device=piv.device)[None, None, :]
    on_diag = free[:, :, None] & (order[:, :, None] == t_idx)    # (free column f, generator t = its order)
    off_diag = free[:, :, None] & (order[:, :, None] != t_idx)
    eq_den = (gens == den).all(dim=-1)
    is_zero = (gens == 0).all(dim=-1)
    assert bool((eq_den | ~on_diag)[ok].all()) and bool((is_zero | ~off_diag)[ok].all())
    nfree = free.sum(dim=1)
    unused = t_idx[0] >= nfree[:, None, None].clamp(max=G)       # [B, 1, G] broadcast over rows
    assert bool((is_zero | ~unused.expand(-1, nvars, -1))[ok].all())


def test_c3_full_size_properties(eng):
    """BASELINE.json configs[2] at its full size (2^18 systems): A * particular == den * b for the consistent
    half, A * generators == 0 everywhere, every odd (random right-hand side) system of rank-10 A is inconsistent or
    satisfies the equations; all checked modulo a prime outside the table, on the device."""
    import torch
    import bench
    q = 1048573
    data = bench.make_inputs(16, 1 << 18, 20260003, "c3")
    A = torch.from_numpy(data["A"]).cuda()
    b = torch.from_numpy(data["b"]).cuda()
    res = eng.solve_batch(A, b, a_abs_max=250, b_abs_max=int(np.abs(data["b"]).max()), max_rank=10, gen_cap=6)
    st = res.status
    assert int((st & ~(2 | 32)).count_nonzero()) == 0
    ok = (st & 2) == 0
    assert bool(ok[0::2].all())                                  # b = A x0 is consistent
    assert int(ok[1::2].count_nonzero()) < 64                    # random b: inconsistent with overwhelming probability
    assert bool((res.rank == 10).all())
    Aq = (A.to(torch.float64) % q)
    part = _limbs_mod(res.particular, q)                         # [B, 16]
    den = _limbs_mod(res.den, q)                                 # [B]
    gens = _limbs_mod(res.generators, q)                         # [B, 16, 6]
    lhs = torch.einsum("bij,bj->bi", Aq, part) % q
    rhs = (den[:, None] * (b.to(torch.float64) % q)) % q
    assert bool((lhs[ok] == rhs[ok]).all())
    assert bool(((torch.einsum("bij,bjk->bik", Aq, gens) % q)[ok] == 0).all())
    assert bool((res.den != 0).any(dim=-1)[ok].all())            # exact words: a non-zero integer may vanish modulo q
    _assert_canonical_solution_form(res, 16, ok)                 # free variables zero, gen[f] = den, unused columns zero


def test_c4_full_size_properties(eng):
    """BASELINE.json configs[3] at its full size (2^16 matrices 64 x 64, 11-limb results): A * adj == det * I modulo
    a prime outside the table for every matrix, in chunks of 2^13, on the device."""
    import torch
    q = 1048573
    rng = np.random.Generator(np.random.PCG64(20260004))
    eye = torch.eye(64, dtype=torch.float64, device="cuda")
    nsing = 0
    for chunk in range(8):
        A = torch.from_numpy(rng.integers(-5, 6, size=(1 << 13, 64, 64), dtype=np.int32)).cuda()
        res = eng.inverse_batch(A, a_abs_max=5)
        assert res.plan.limbs == 11 and int((res.status & ~(1 | 32)).count_nonzero()) == 0
        adj = _limbs_mod(res.adj, q)
        det = _limbs_mod(res.det, q)
        prod = torch.matmul(A.to(torch.float64) % q, adj) % q
        want = (det[:, None, None] * eye[None]) % q
        sing = (res.status & 1) != 0
        nsing += int(sing.count_nonzero())
        assert bool((prod[~sing] == want[~sing]).all())
        del A, res, adj, det, prod, want
    assert nsing < 8                                            # random 64 x 64 integer matrices are almost never singular


def test_c4_kernel_basis_full_size_properties(eng):
    """BASELINE.json configs[3], second half, at its full size: 2^16 products B(64x48) C(48x64); the device kernel
    basis has 16 generators with A G == 0 and den on the free positions, modulo an independent prime."""
    import torch
    import bench
    q = 1048573
    for chunk in range(16):
        data = bench.make_inputs(64, 1 << 12, 20260040 + chunk, "c4ker")
        A = torch.from_numpy(data["A"]).cuda()
        z = torch.zeros((1 << 12, 64), dtype=torch.int32, device="cuda")
        res = eng.solve_batch(A, z, a_abs_max=int(np.abs(data["A"]).max()), b_abs_max=0, max_rank=48, gen_cap=16)
        assert int((res.status & ~32).count_nonzero()) == 0 and bool((res.rank == 48).all())
        gens = _limbs_mod(res.generators, q)                     # [B, 64, 16]
        den = _limbs_mod(res.den, q)
        assert bool(((torch.matmul(A.to(torch.float64) % q, gens) % q) == 0).all())
        assert int(_limbs_mod(res.particular, q).count_nonzero()) == 0
        # den is a non-zero integer and every generator column has a non-zero entry (exact words, not residues:
        # a non-zero integer may vanish modulo q)
        assert bool((res.den != 0).any(dim=-1).all())
        assert bool((res.generators != 0).any(dim=-1).any(dim=1).all())
        _assert_canonical_solution_form(res, 64)
        del A, res, gens, den


# ------------------------------------------------------------------ config 3: 16x17 rank-10 systems
def test_c3_golden(eng):
    g = golden_io.load("c3_16x17")
    assert check_systems(eng, g["cases"]) > 100
    assert check_systems(eng, g["cases"][:128], max_rank=10) > 10


# ------------------------------------------------------------------ config 4: 64x64
def test_c4_golden(eng):
    g = golden_io.load("c4_64x64")
    inv = g["inverse_cases"]
    A = np_batch([c["A"] for c in inv])
    res = eng.inverse_batch(A, a_abs_max=5)
    assert res.plan.n_primes == 12 and res.plan.limbs == 11
    adj, det = limbs_to_ints(res.adj), limbs_to_ints(res.det)
    for i, c in enumerate(inv):
        assert int(res.status[i]) == 0
        pq = pq_grid_from_num(adj[i], det[i])
        assert golden_io.digest_pq(pq) == c["inverse_sha"]
        assert pq[:64] == c["inverse_row0"]
        assert det[i] == ref_port.bareiss_det(c["A"])
    ker = g["kernel_cases"]
    K = np_batch([c["A"] for c in ker])
    z = np.zeros((len(ker), 64), dtype=np.int32)
    amax = int(np.abs(K).max())
    sres = eng.solve_batch(K, z, a_abs_max=amax, b_abs_max=0, max_rank=48, gen_cap=16)
    for i, c in enumerate(ker):
        got = affine_pq(sres, i, 64, "default")
        assert got["status"] == c["status"] and got["gen_cols"] == c["gen_cols"] and got["sha"] == c["sha"]


# ------------------------------------------------------------------ config 5 stand-ins and CRT
def test_c5_standins_by_prime(eng):
    g = golden_io.load("c5_standins")
    for c in g["cases"]:
        n = c["n"]
        if n > 200:
            continue
        rng = np.random.Generator(np.random.PCG64(c["seed"]))
        A = rng.integers(-5, 6, size=(n, n), dtype=np.int64).astype(np.int32)
        K, bits = eng.det_large_prime_count(n, 5)
        half = K // 2
        r0 = eng.det_large_residues(A, 0, half)                 # two "ranks" worth of primes
        r1 = eng.det_large_residues(A, half, K - half)
        res = np.concatenate([r0, r1])
        primes = eng.primes(K)
        want = int(c["det"])
        assert [int(x) for x in res] == [want % int(p) for p in primes]
        limbs = int(bits + 2) // 32 + 1
        words = eng.crt_signed(res, limbs)
        assert limbs_to_ints(words) == want


def test_crt_signed_many_primes(eng):
    import random
    rnd = random.Random(5)
    for K in (1, 2, 3, 40, 97, 300):
        primes = [int(p) for p in eng.primes(K)]
        M = 1
        for p in primes:
            M *= p
        for x in (0, 1, -1, M // 2, -(M // 2), rnd.randrange(-(M // 2), M // 2)):
            res = np.array([x % p for p in primes], dtype=np.uint32)
            limbs = K + 1
            assert limbs_to_ints(eng.crt_signed(res, limbs)) == x


# ------------------------------------------------------------------ bad primes and error paths
def test_bad_prime_is_detected_and_replaced(eng):
    p0, p1 = (int(p) for p in eng.primes(2))
    # the first pivot candidate is divisible by the first prime: that prime picks another row
    mats = [[[p0, 1, 0], [1, 1, 0], [0, 0, 1]], [[p0, 2, 1], [3, p1, 1], [1, 1, 1]], [[1, 2, 3], [4, 5, 6], [7, 8, 10]]]
    res = eng.rref_batch(np_batch(mats), 3, a_abs_max=2**31 - 1, b_abs_max=0)
    assert int(res.status[0]) & 32 and int(res.status[1]) & 32      # informational "retried" flag
    assert not np.any(res.status & ~32)
    num, den = limbs_to_ints(res.num), limbs_to_ints(res.den)
    for i, A in enumerate(mats):
        R, piv = ref_port.row_reduce(A, 3)
        assert pq_grid_from_num(num[i], den[i]) == pq_of_fracs(R)
        assert den[i] == ref_port.bareiss_det(A)
    d = eng.det_batch(np_batch(mats), a_abs_max=2**31 - 1)
    assert limbs_to_ints(d.det) == [ref_port.bareiss_det(A) for A in mats]


def test_declared_bound_violation_is_flagged(eng):
    res = eng.det_batch(np_batch([[[1, 2], [3, 4]], [[100, 2], [3, 4]]]), a_abs_max=5)
    assert int(res.status[0]) == 0 and int(res.status[1]) & 4


def test_shape_errors(eng):
    from linalg_solver_b200 import LsxError
    with pytest.raises(ValueError):
        eng.inverse_batch(np.zeros((1, 2, 3), dtype=np.int32))
    with pytest.raises(LsxError):
        eng.rref_batch(np.zeros((1, 2, 3), dtype=np.int32), 4)
    r = eng.rank_batch(np.zeros((0, 3, 3), dtype=np.int32))
    assert r.rank.shape == (0,)


def test_random_shapes_against_oracle(eng):
    import random
    rnd = random.Random(11)
    for _ in range(12):
        m, n = rnd.randint(1, 12), rnd.randint(1, 14)
        mats = []
        for _ in range(24):
            r = rnd.randint(0, min(m, n))
            B = [[rnd.randint(-5, 5) for _ in range(r)] for _ in range(m)]
            C = [[rnd.randint(-5, 5) for _ in range(n)] for _ in range(r)]
            mats.append([[sum(B[i][k] * C[k][j] for k in range(r)) for j in range(n)] for i in range(m)])
        check_rref_against_oracle(eng, mats, rnd.randint(1, n))


def test_device_memory_path_matches_host_path(eng):
    import torch
    rng = np.random.Generator(np.random.PCG64(3))
    A = rng.integers(-5, 6, size=(4096, 8, 8), dtype=np.int32)
    host = eng.inverse_batch(A, a_abs_max=5)
    dev = eng.inverse_batch(torch.from_numpy(A).cuda(), a_abs_max=5)
    torch.cuda.synchronize()
    assert np.array_equal(host.adj.view(np.int32), dev.adj.cpu().numpy().reshape(host.adj.shape))
    assert np.array_equal(host.det.view(np.int32), dev.det.cpu().numpy().reshape(host.det.shape))
    assert np.array_equal(host.status, dev.status.cpu().numpy())


def test_fused_small_inverse_equals_tile_path(eng, monkeypatch):
    """The fused register-resident kernel in its two arithmetics -- exact 32-bit integers (Bareiss, the default) and
    residues modulo one prime (LSX_TPM_ALGO=mont, with and without the plain-integer head) -- and the multi-prime
    shared-memory path must agree word for word, for every n <= 8, including singular inputs, row swaps and ragged
    batch sizes."""
    rng = np.random.Generator(np.random.PCG64(99))
    for n in range(1, 9):
        for B in (1, 127, 128, 129, 1000):
            A = rng.integers(-5, 6, size=(B, n, n), dtype=np.int32)
            A[::7, :, 0] = 0 if n == 1 else A[::7, :, 0] * (rng.integers(0, 2, size=(len(A[::7]), n)))   # zeros in column 0
            if n > 1:
                A[::11, n - 1] = A[::11, 0]                    # singular
            fused = eng.inverse_batch(A, a_abs_max=5)
            monkeypatch.setenv("LSX_TPM_ALGO", "mont")          # the residue kernel
            mont = eng.inverse_batch(A, a_abs_max=5)
            monkeypatch.setenv("LSX_TPM_HEAD", "0")             # ... and without its plain-integer head steps
            nohead = eng.inverse_batch(A, a_abs_max=5)
            monkeypatch.delenv("LSX_TPM_HEAD")
            monkeypatch.delenv("LSX_TPM_ALGO")
            for alt in (mont, nohead):
                assert np.array_equal(fused.adj, alt.adj) and np.array_equal(fused.det, alt.det)
                assert np.array_equal(fused.status, alt.status)
            monkeypatch.setenv("LSX_DISABLE_SMALL", "1")
            tile = eng.inverse_batch(A, a_abs_max=5)
            monkeypatch.delenv("LSX_DISABLE_SMALL")
            assert np.array_equal(fused.status, tile.status & ~32)
            assert np.array_equal(fused.det, tile.det)
            assert np.array_equal(fused.adj, tile.adj)
    # magnitudes where the 32-bit steps of the integer kernel do not apply (its all-64-bit instantiation) and minors
    # close to 2^31: still the fused path (one limb), still equal to the other two
    for n, amax in ((2, 30000), (3, 600), (4, 100), (5, 30), (6, 14), (7, 8)):
        A = rng.integers(-amax, amax + 1, size=(777, n, n), dtype=np.int32)
        A[::13, 0] = 0
        A[::13, 0, n - 1] = amax
        A[::17, n - 1] = A[::17, 0]
        fused = eng.inverse_batch(A, a_abs_max=amax)
        assert fused.plan.limbs == 1
        monkeypatch.setenv("LSX_TPM_ALGO", "mont")
        mont = eng.inverse_batch(A, a_abs_max=amax)
        monkeypatch.delenv("LSX_TPM_ALGO")
        monkeypatch.setenv("LSX_DISABLE_SMALL", "1")
        tile = eng.inverse_batch(A, a_abs_max=amax)
        monkeypatch.delenv("LSX_DISABLE_SMALL")
        for alt in (mont, tile):
            assert np.array_equal(fused.adj, alt.adj) and np.array_equal(fused.det, alt.det), (n, amax)
            assert np.array_equal(fused.status, alt.status & ~32), (n, amax)
        for i in range(0, 777, 97):
            inv = ref_port.inverse(A[i].tolist())
            d = int(fused.det.view(np.int32).reshape(-1)[i])
            assert d == ref_port.bareiss_det(A[i].tolist())
            if inv is not None:
                adj = fused.adj.view(np.int32).reshape(777, n, n)[i]
                assert [[Fraction(int(x), d) for x in row] for row in adj] == inv
    # entries above the declared bound are flagged, not mis-computed
    A = rng.integers(-5, 6, size=(300, 8, 8), dtype=np.int32)
    A[5, 3, 3] = 77
    r = eng.inverse_batch(A, a_abs_max=5)
    assert int(r.status[5]) & 4 and not np.any(np.delete(r.status, 5) & 4)
    # magnitudes the single-prime argument does not cover fall back to the multi-prime path and stay exact
    A = rng.integers(-1000, 1001, size=(64, 8, 8), dtype=np.int32)
    r = eng.inverse_batch(A)
    assert r.plan.limbs > 1
    adj, det = limbs_to_ints(r.adj), limbs_to_ints(r.det)
    for i in range(0, 64, 9):
        inv = ref_port.inverse(A[i].tolist())
        assert [[Fraction(x, det[i]) for x in row] for row in adj[i]] == inv


def test_blocked_lu_residues_match_tile_path(eng, monkeypatch):
    """The global-memory blocked LU (used above n ~ 220) against the shared-memory tile kernel on
    sizes both can do, including ragged n, zero pivots (row swaps) and singular matrices."""
    rng = np.random.Generator(np.random.PCG64(77))
    for n in (1, 5, 63, 64, 65, 100, 130, 200):
        A = rng.integers(-5, 6, size=(n, n), dtype=np.int32)
        variants = [A]
        if n > 2:
            B = A.copy()
            B[0, 0] = 0
            B[1, 0] = 0
            B[n // 2, n // 2:] = 0                       # forces later swaps too
            variants.append(B)
            C = A.copy()
            C[n - 1] = C[0]                              # singular
            variants.append(C)
        for M in variants:
            tile = eng.det_large_residues(M, 3, 9)
            monkeypatch.setenv("LSX_FORCE_BLOCKED", "1")
            blk = eng.det_large_residues(M, 3, 9)
            monkeypatch.delenv("LSX_FORCE_BLOCKED")
            assert np.array_equal(tile, blk), n
    assert not np.any(eng.det_large_residues(C, 0, 4))


def test_blocked_lu_tensor_core_path_vs_numpy_oracle(eng, monkeypatch):
    """Recursive blocked LU with the tcgen05 int8-split trailing update (depths 64/128/256) against
    oracle/det_mod_p.py and against the same LU with the integer-pipe GEMM (LSX_NO_TC), on sizes that
    exercise full and ragged outer blocks, row swaps and singular inputs."""
    from oracle.det_mod_p import det_mod_p
    rng = np.random.Generator(np.random.PCG64(505))
    for n in (230, 257, 320, 520, 1000):
        A = rng.integers(-5, 6, size=(n, n), dtype=np.int32)
        B = A.copy()
        B[0, 0] = 0
        B[1, 0] = 0
        B[n // 2, : n // 2 + 3] = 0                          # zero pivot deep inside: swap across blocks
        C = A.copy()
        C[n - 2] = C[3]                                      # singular
        for M in (A, B, C):
            tc = eng.det_large_residues(M, 5, 3)
            monkeypatch.setenv("LSX_NO_TC", "1")
            ip = eng.det_large_residues(M, 5, 3)
            monkeypatch.delenv("LSX_NO_TC")
            assert np.array_equal(tc, ip), n
            if n <= 520:
                primes = eng.primes(8)[5:8]
                assert [int(x) for x in tc] == [det_mod_p(M, int(p)) for p in primes], n
        assert not np.any(eng.det_large_residues(C, 0, 2))
        assert int(eng.det_large_residues(A, 7, 1)[0]) == det_mod_p(A, int(eng.primes(8)[7])) or n > 520   # one prime


def test_c5_standin_256_blocked_and_sharded_crt(eng):
    from linalg_solver_b200 import dist as lsx_dist
    g = golden_io.load("c5_standins")
    c = [x for x in g["cases"] if x["n"] == 256][0]
    rng = np.random.Generator(np.random.PCG64(c["seed"]))
    A = rng.integers(-5, 6, size=(256, 256), dtype=np.int64).astype(np.int32)
    words, K = lsx_dist.det_large_sharded(eng, A, 5)
    assert limbs_to_ints(words) == int(c["det"])
    assert K == eng.det_large_prime_count(256, 5)[0]


@pytest.mark.parametrize("n", [384, 512])
def test_c5_standins_blocked_tensor_path_exact(eng, n):
    """Full config 5 route (blocked LU with tcgen05 updates -> residues -> CRT) against the exact DomainMatrix
    determinants of the 384 and 512 stand-ins."""
    from linalg_solver_b200 import dist as lsx_dist
    g = golden_io.load("c5_standins")
    c = [x for x in g["cases"] if x["n"] == n][0]
    rng = np.random.Generator(np.random.PCG64(c["seed"]))
    A = rng.integers(-5, 6, size=(n, n), dtype=np.int64).astype(np.int32)
    words, K = lsx_dist.det_large_sharded(eng, A, 5)
    assert limbs_to_ints(words) == int(c["det"])
    # prime count from the matrix's own row/column norms: fewer primes, same determinant
    words2, K2 = lsx_dist.det_large_sharded(eng, A)
    assert K2 < K and K2 == eng.det_large_prime_count_for(torch_or_np(A))[0]
    assert limbs_to_ints(words2) == int(c["det"])
    Z = A.copy()
    Z[:, 7] = 0
    assert eng.det_large_prime_count_for(Z) == (1, 0.0)


def test_random_differential_against_the_oracle(eng):
    """Randomised differential test: shapes up to 12 x 14, entries up to +-60, every rank from 0 to full, random
    bar_col; rref, rank, determinant, inverse and find_preimage_of through the C-ABI against oracle/ref_port.py
    (itself pinned to the reference's goldens)."""
    import random
    rnd = random.Random(20260018)
    rng = np.random.Generator(np.random.PCG64(20260018))
    for trial in range(60):
        m, n = rnd.randint(1, 12), rnd.randint(1, 14)
        B = 7
        amp = rnd.choice([2, 5, 60])
        mats = np.zeros((B, m, n), dtype=np.int32)
        for i in range(B):
            rk = rnd.randint(0, min(m, n))
            if rk == min(m, n) or rk == 0:
                mats[i] = rng.integers(-amp, amp + 1, size=(m, n)) if rk else 0
            else:
                mats[i] = rng.integers(-3, 4, size=(m, rk)) @ rng.integers(-3, 4, size=(rk, n))
        bar = rnd.randint(1, n)
        res = eng.rref_batch(mats, bar)
        num, den = limbs_to_ints(res.num), limbs_to_ints(res.den)
        rks = eng.rank_batch(mats).rank
        for i in range(B):
            R, piv = ref_port.row_reduce(mats[i].tolist(), bar)
            assert int(res.status[i]) & ~32 == 0
            assert [(k, int(res.pivot_col[i][k])) for k in range(int(res.rank[i]))] == piv, (m, n, bar)
            assert pq_grid_from_num(num[i], den[i]) == pq_of_fracs(R), (m, n, bar, mats[i].tolist())
            assert int(rks[i]) == ref_port.rank(mats[i].tolist())
        if m == n:
            inv = eng.inverse_batch(mats)
            adj, det = limbs_to_ints(inv.adj), limbs_to_ints(inv.det)
            dets = limbs_to_ints(eng.det_batch(mats).det)
            for i in range(B):
                want = ref_port.inverse(mats[i].tolist())
                assert dets[i] == det[i] == ref_port.bareiss_det(mats[i].tolist())
                if want is None:
                    assert int(inv.status[i]) & 1
                else:
                    assert pq_grid_from_num(adj[i], det[i]) == pq_of_fracs(want)
        b = rng.integers(-amp, amp + 1, size=(B, m), dtype=np.int32)
        b[::2] = np.einsum("bij,bj->bi", mats[::2].astype(np.int64), rng.integers(-3, 4, size=(len(mats[::2]), n))).astype(np.int32)
        sol = eng.solve_batch(mats, b)
        for i in range(B):
            want = ref_port.find_preimage_of(mats[i].tolist(), b[i].tolist())
            got = affine_pq(sol, i, n, "logged")
            if want is None:
                assert got["status"] == "nosolution", (m, n)
            else:
                part, gens = want
                assert got["status"] == "ok" and got["particular"] == [[x.numerator, x.denominator] for x in part]
                k = n - int(sol.rank[i])
                flat = [[gens[t][r].numerator, gens[t][r].denominator] for r in range(n) for t in range(k)] if k else []
                assert got["generators"] == flat


def test_rref_trace_device_memory_equals_host_memory(eng):
    """lsx_rref_trace with LSX_MEM_DEVICE buffers (enqueue only) against the LSX_MEM_HOST call: same op log, frames,
    step counts and pivot columns for every prime."""
    import ctypes
    import torch
    from linalg_solver_b200 import _lib
    rng = np.random.Generator(np.random.PCG64(12))
    for m, n, bar in [(3, 4, 3), (6, 7, 6), (5, 9, 4), (12, 12, 12)]:
        A = rng.integers(-5, 6, size=(m, n), dtype=np.int32)
        A[0, 0] = 0
        K = 5
        max_ops = _lib.lib.lsx_rref_trace_max_ops(m, n, bar)
        slots = min(m, bar)
        h_ops = np.zeros((K, max_ops, 4), np.int32)
        h_fr = np.zeros((K, max_ops, m, n), np.uint32)
        h_n = np.zeros(K, np.int32)
        h_pc = np.zeros((K, slots), np.int32)
        ctx = eng._ctx
        eng.set_stream(None)
        assert _lib.lib.lsx_rref_trace(ctx, A.ctypes.data, m, n, bar, K, _lib.MEM_HOST, h_ops.ctypes.data, h_fr.ctypes.data,
                                       h_n.ctypes.data, h_pc.ctypes.data) == 0
        dA = torch.from_numpy(A).cuda()
        d_ops = torch.zeros((K, max_ops, 4), dtype=torch.int32, device="cuda")
        d_fr = torch.zeros((K, max_ops, m, n), dtype=torch.int32, device="cuda")
        d_n = torch.zeros(K, dtype=torch.int32, device="cuda")
        d_pc = torch.zeros((K, slots), dtype=torch.int32, device="cuda")
        eng.set_stream(torch.cuda.current_stream().cuda_stream or 1)
        assert _lib.lib.lsx_rref_trace(ctx, dA.data_ptr(), m, n, bar, K, _lib.MEM_DEVICE, d_ops.data_ptr(), d_fr.data_ptr(),
                                       d_n.data_ptr(), d_pc.data_ptr()) == 0
        torch.cuda.synchronize()
        T = int(h_n[0])
        assert T > 0 and np.array_equal(d_n.cpu().numpy(), h_n) and np.array_equal(d_pc.cpu().numpy(), h_pc)
        assert np.array_equal(d_ops.cpu().numpy()[:, :T], h_ops[:, :T])
        assert np.array_equal(d_fr.cpu().numpy().view(np.uint32)[:, :T], h_fr[:, :T])
    assert _lib.lib.lsx_rref_trace_max_ops(3, 4, 9) < 0          # bar_col beyond n


def test_inverse_int8_input_container(eng):
    """lsx_inverse_batch_i8: the same matrices in an int8 container give the same words as int32, from host and
    from device memory, for every size of the fused kernel; larger sizes are widened by the Python layer."""
    import torch
    rng = np.random.Generator(np.random.PCG64(88))
    for n in (1, 2, 3, 5, 7, 8):
        A = rng.integers(-5, 6, size=(301, n, n), dtype=np.int32)
        A[7] = 0                                              # singular
        want = eng.inverse_batch(A, a_abs_max=5)
        got = eng.inverse_batch(A.astype(np.int8), a_abs_max=5)
        for k in ("adj", "det", "status"):
            assert np.array_equal(getattr(want, k), getattr(got, k)), (n, k)
        dev = eng.inverse_batch(torch.from_numpy(A.astype(np.int8)).cuda(), a_abs_max=5)
        assert np.array_equal(dev.adj.cpu().numpy().view(np.uint32), want.adj)       # torch holds the words as int32
        assert np.array_equal(dev.status.cpu().numpy(), want.status)
    B = rng.integers(-5, 6, size=(9, 16, 16), dtype=np.int32)
    w16 = eng.inverse_batch(B, a_abs_max=5)
    g16 = eng.inverse_batch(B.astype(np.int8), a_abs_max=5)
    assert np.array_equal(w16.adj, g16.adj) and np.array_equal(w16.det, g16.det)


def _planted(n, seed):
    """A = L U with unit-lower L and upper U (entries in {-1, 0, 1}, diagonal of U in +-{1, 2, 3}): the determinant
    is the product of U's diagonal (planted matrices with known determinant).  The product
    runs in float64 BLAS, exact because every entry of A is below n."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = np.tril(rng.integers(-1, 2, size=(n, n)), -1).astype(np.float64) + np.eye(n)
    U = np.triu(rng.integers(-1, 2, size=(n, n)), 1).astype(np.float64)
    d = rng.choice(np.array([-3, -2, -1, 1, 2, 3]), size=n)
    U[np.arange(n), np.arange(n)] = d
    A = L @ U
    assert np.abs(A).max() < 2**31
    det = 1
    for x in d.tolist():
        det *= int(x)
    return A.astype(np.int32), det


def test_c5_full_size_planted_determinant_exact(eng):
    """BASELINE.json configs[4] at its full size, 4096 x 4096: exact determinant of a planted matrix through the
    blocked tensor-core LU, the data-dependent Hadamard bound and the CRT."""
    from linalg_solver_b200 import dist as lsx_dist
    A, det = _planted(4096, 4096)
    words, K = lsx_dist.det_large_sharded(eng, torch_or_np(A))
    assert K > 200
    assert limbs_to_ints(words.cpu().numpy()) == det


def test_c5_full_size_matrix_against_cpu_residues(eng):
    """BASELINE.json configs[4], the very matrix bench.py times (4096 x 4096, PCG64(20260005), entries in [-5, 5]):
    det mod p for two table primes OUTSIDE the CRT set, computed on the CPU by oracle/det_mod_p.py (numpy int64
    elimination, ~11 minutes per prime; oracle/gen_c5_residues.py, stored in tests/golden/c5_full_residues), against
    (i) the device residues for those primes and (ii) the exact CRT determinant reduced modulo them."""
    from linalg_solver_b200 import dist as lsx_dist
    g = golden_io.load("c5_full_residues")
    rng = np.random.Generator(np.random.PCG64(g["seed"]))
    A = torch_or_np(rng.integers(-5, 6, size=(g["n"], g["n"]), dtype=np.int32))
    words, K = lsx_dist.det_large_sharded(eng, A)
    det = limbs_to_ints(words.cpu().numpy())
    for e in g["entries"]:
        assert K <= e["prime_index"]                          # an independent prime: not part of the lift
        assert int(eng.primes(e["prime_index"] + 1)[-1]) == e["prime"]
        got = eng.det_large_residues(A, e["prime_index"], 1).cpu().numpy().view(np.uint32)
        assert int(got[0]) == e["residue"]
        assert det % e["prime"] == e["residue"]


def test_blocked_lu_beyond_the_register_panel(eng):
    """More than 4096 rows: the first base panels take the global-memory fallback (k_panel_gmem); residues of a
    planted matrix for three primes, n not a multiple of 8."""
    n = 4100 + 3
    A, det = _planted(n, 7)
    res = eng.det_large_residues(torch_or_np(A), 2, 3).cpu().numpy()
    primes = eng.primes(5)[2:5]
    assert [int(x) for x in res] == [det % int(p) for p in primes]


def torch_or_np(A):
    import torch
    return torch.from_numpy(A).cuda()


def test_subwarp_kernel_equals_tile_path(eng, monkeypatch):
    """Fused sub-warp kernel (row per lane, all primes + CRT in one launch) against the tile path, word
    for word, over every lane-group size, ragged shapes, rank-deficient inputs and all operations."""
    import random
    rnd = random.Random(23)
    rng = np.random.Generator(np.random.PCG64(23))
    shapes = [(1, 1), (2, 3), (4, 4), (4, 5), (3, 9), (5, 5), (8, 9), (7, 12), (8, 16), (9, 9), (16, 17), (13, 20),
              (16, 32), (17, 17), (32, 33), (20, 7)]
    for m, n in shapes:
        B = 67
        rk = rng.integers(0, min(m, n) + 1, size=B)
        mats = np.zeros((B, m, n), dtype=np.int32)
        for i in range(B):
            L_ = rng.integers(-3, 4, size=(m, rk[i]))
            R_ = rng.integers(-3, 4, size=(rk[i], n))
            mats[i] = L_ @ R_
        mats[::5] = rng.integers(-9, 10, size=mats[::5].shape)
        bar = rnd.randint(1, min(n, 32))
        def both(f):
            x = f()
            monkeypatch.setenv("LSX_DISABLE_SUBWARP", "1")
            y = f()
            monkeypatch.delenv("LSX_DISABLE_SUBWARP")
            return x, y
        x, y = both(lambda: eng.rref_batch(mats, bar))
        assert np.array_equal(x.status & ~32, y.status & ~32), (m, n, bar)
        for f in ("num", "den", "pivot_col", "rank"):
            assert np.array_equal(getattr(x, f), getattr(y, f)), (m, n, bar, f)
        if n <= 32:
            x, y = both(lambda: eng.rank_batch(mats))
            assert np.array_equal(x.rank, y.rank)
            b = rng.integers(-9, 10, size=(B, m), dtype=np.int32)
            b[::2] = np.einsum("bij,bj->bi", mats[::2], rng.integers(-2, 3, size=(len(mats[::2]), n)))
            x, y = both(lambda: eng.solve_batch(mats, b))
            assert np.array_equal(x.status & ~32, y.status & ~32)
            ok = (x.status & 2) == 0
            for f in ("den", "particular", "generators", "pivot_col", "rank"):
                assert np.array_equal(getattr(x, f)[ok], getattr(y, f)[ok]), (m, n, f)
        if m == n:
            x, y = both(lambda: eng.det_batch(mats))
            assert np.array_equal(x.det, y.det) and np.array_equal(x.rank, y.rank)
            if n <= 16:
                big = mats * 1000 + 1                        # magnitudes beyond the single-prime kernel
                x, y = both(lambda: eng.inverse_batch(big))
                assert np.array_equal(x.status & ~32, y.status & ~32)
                assert np.array_equal(x.adj, y.adj) and np.array_equal(x.det, y.det)


def test_register_tiled_kernel_equals_smem_kernel(eng, monkeypatch):
    """k_tile_reg (tile in registers, logical row permutation, finished column blocks skipped) against
    k_tile_elim, word for word, on mid-size shapes with rank-deficient inputs and forced row swaps."""
    rng = np.random.Generator(np.random.PCG64(41))
    def both(f):
        x = f()
        monkeypatch.setenv("LSX_DISABLE_TILE_REG", "1")
        y = f()
        monkeypatch.delenv("LSX_DISABLE_TILE_REG")
        return x, y
    for m, n, bar in [(34, 34, 34), (48, 48, 48), (40, 41, 40), (33, 70, 33), (64, 64, 64), (64, 65, 64), (50, 100, 50),
                      (64, 128, 64), (100, 90, 60), (80, 80, 80), (20, 40, 15)]:
        B = 9
        mats = np.zeros((B, m, n), dtype=np.int32)
        for i in range(B):
            rk = int(rng.integers(max(1, min(m, n) - 20), min(m, n) + 1))
            mats[i] = rng.integers(-2, 3, size=(m, rk)) @ rng.integers(-2, 3, size=(rk, n))
        mats[0] = rng.integers(-5, 6, size=(m, n))
        mats[1] = rng.integers(-5, 6, size=(m, n))
        mats[1, 0, 0] = 0
        mats[1, 1, :2] = 0                                   # swaps in the first columns
        x, y = both(lambda: eng.rref_batch(mats, bar))
        assert np.array_equal(x.status, y.status), (m, n)
        for f in ("num", "den", "pivot_col", "rank"):
            assert np.array_equal(getattr(x, f), getattr(y, f)), (m, n, bar, f)
        x, y = both(lambda: eng.rank_batch(mats))
        assert np.array_equal(x.rank, y.rank)
        if m == n:
            x, y = both(lambda: eng.det_batch(mats))
            assert np.array_equal(x.det, y.det)
            if n <= 64:
                x, y = both(lambda: eng.inverse_batch(mats))
                assert np.array_equal(x.status, y.status)
                assert np.array_equal(x.adj, y.adj) and np.array_equal(x.det, y.det)
        if bar == n or True:
            b = rng.integers(-5, 6, size=(B, m), dtype=np.int32)
            b[::2] = np.einsum("bij,bj->bi", mats[::2], rng.integers(-2, 3, size=(len(mats[::2]), n)))
            x, y = both(lambda: eng.solve_batch(mats, b))
            assert np.array_equal(x.status, y.status)
            ok = (x.status & 2) == 0
            for f in ("den", "particular", "generators", "pivot_col", "rank"):
                assert np.array_equal(getattr(x, f)[ok], getattr(y, f)[ok]), (m, n, f)
    # an exact check against the oracle on one rank-deficient mid-size input
    A = (rng.integers(-2, 3, size=(36, 30)) @ rng.integers(-2, 3, size=(30, 40))).astype(np.int32)
    check_rref_against_oracle(eng, [A.tolist()], 38)


def test_device_calls_capture_into_a_cuda_graph(eng):
    """Device-memory calls only enqueue on the caller's stream, so a sequence of them is capturable into one CUDA graph
    (Engine.capture): the fused sub-warp kernel (config 1's three calls, a solve with its output memsets), the fused
    8x8 kernel and the tile path with its row-bound pass and retry pass.  A replay into zeroed outputs gives the words
    of the direct calls, and the launch count keeps counting."""
    import dataclasses
    import torch
    rng = np.random.Generator(np.random.PCG64(53))
    def tensors(r):
        return [getattr(r, f.name) for f in dataclasses.fields(r) if torch.is_tensor(getattr(r, f.name))]
    A4 = torch.from_numpy(rng.integers(-5, 6, size=(1000, 4, 4), dtype=np.int32)).cuda()
    A8 = torch.from_numpy(rng.integers(-5, 6, size=(512, 8, 8), dtype=np.int32)).cuda()
    A20 = torch.from_numpy(rng.integers(-5, 6, size=(40, 20, 20), dtype=np.int32)).cuda()
    S = torch.from_numpy(rng.integers(-5, 6, size=(300, 6, 9), dtype=np.int32)).cuda()
    b = torch.from_numpy(rng.integers(-5, 6, size=(300, 6), dtype=np.int32)).cuda()
    plans = (eng.plan_det(4, 5), eng.plan_rank(4, 4, 5), eng.plan_rref(4, 4, 3, 5, 5), eng.plan_inverse(8, 5),
             eng.plan_inverse(20, 5), eng.plan_solve(6, 9, 5, 5, 0, 4))
    res = [eng.det_batch(A4, plan=plans[0]), eng.rank_batch(A4, plan=plans[1]), eng.rref_batch(A4, 3, plan=plans[2]),
           eng.inverse_batch(A8, plan=plans[3]), eng.inverse_batch(A20, plan=plans[4]), eng.solve_batch(S, b, plan=plans[5])]
    def step():
        eng.det_batch(A4, plan=plans[0], out=res[0])
        eng.rank_batch(A4, plan=plans[1], out=res[1])
        eng.rref_batch(A4, 3, plan=plans[2], out=res[2])
        eng.inverse_batch(A8, plan=plans[3], out=res[3])
        eng.inverse_batch(A20, plan=plans[4], out=res[4])
        eng.solve_batch(S, b, plan=plans[5], out=res[5])
    step()
    torch.cuda.synchronize()
    want = [t.clone() for r in res for t in tensors(r)]
    cap = eng.capture(step)
    assert cap.kernels >= 6
    for _ in range(2):
        for r in res:
            for t in tensors(r):
                t.zero_()
        n0 = eng.launch_count
        cap.replay()
        torch.cuda.synchronize()
        assert eng.launch_count - n0 == cap.kernels
        got = [t for r in res for t in tensors(r)]
        assert len(got) == len(want) and all(torch.equal(x, y) for x, y in zip(want, got))


def test_solve_leaves_zeros_in_its_unused_slots(eng, monkeypatch):
    """A solve owes zeros in the slots it does not fill (particular solution at free variables, generator columns beyond
    the kernel's dimension, a free variable's row outside its own generator): outputs prefilled with 0xFFFFFFFF come
    back identical to the tile path's, for every group width of the sub-warp kernel, more variables than lanes, fewer
    free variables than generator columns, truncated generators and systems flagged for a declared-magnitude violation
    (all zero).  (Zeroing them inside the sub-warp kernel instead of by memsets in front of it was built and measured
    neutral on config 3 -- +0.06 ms in the kernel, -0.05 ms outside -- and is not in.)"""
    import torch
    rng = np.random.Generator(np.random.PCG64(59))
    for m, nv, rk, gen_cap in ((3, 4, 2, 3), (6, 8, 3, 6), (8, 16, 5, 12), (16, 16, 10, 6), (16, 16, 16, 2), (12, 20, 7, 4),
                               (30, 32, 20, 13), (5, 3, 3, 2)):
        B = 200
        Bm = rng.integers(-3, 4, size=(B, m, rk)); Cm = rng.integers(-3, 4, size=(B, rk, nv))
        A = np.einsum("bik,bkj->bij", Bm, Cm).astype(np.int32)
        b = rng.integers(-5, 6, size=(B, m)).astype(np.int32)
        b[::2] = np.einsum("bij,bj->bi", A[::2], rng.integers(-3, 4, size=(B // 2, nv)))
        amax = int(np.abs(A).max())
        A[7, 0, 0] = amax + 1000                              # flagged: above the declared magnitude
        plan = eng.plan_solve(m, nv, amax, int(np.abs(b).max()), 0, gen_cap)
        At, bt = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
        x = eng.solve_batch(At, bt, plan=plan)
        for f in ("particular", "generators", "den"):
            getattr(x, f).fill_(-1)
        x = eng.solve_batch(At, bt, plan=plan, out=x)
        monkeypatch.setenv("LSX_DISABLE_SUBWARP", "1")
        y = eng.solve_batch(At, bt, plan=plan)
        monkeypatch.delenv("LSX_DISABLE_SUBWARP")
        assert torch.equal(x.status, y.status), (m, nv)
        assert int(x.status[7]) & 4
        ok = ((x.status & 2) == 0).cpu().numpy()
        ok[7] = True                                          # flagged: both paths leave zeros
        assert not np.any(x.particular[7].cpu().numpy()) and not np.any(x.generators[7].cpu().numpy())
        for f in ("particular", "generators", "rank", "pivot_col"):
            assert np.array_equal(getattr(x, f).cpu().numpy()[ok], getattr(y, f).cpu().numpy()[ok]), (m, nv, f)


def test_prime_count_from_row_norms(eng, monkeypatch):
    """The tile path runs only the primes the Hadamard bound of the batch's own row norms needs (k_row_bound), never more
    than the plan: identical words with the data bound switched off, fewer primes on random data, the plan's count on
    worst-case data, and exact answers where the determinant EQUALS the row-norm bound (diagonal matrices)."""
    rng = np.random.Generator(np.random.PCG64(47))
    def both(f):
        x = f()
        kx = eng.last_prime_count()
        monkeypatch.setenv("LSX_NO_DATA_BOUND", "1")
        y = f()
        ky = eng.last_prime_count()
        monkeypatch.delenv("LSX_NO_DATA_BOUND")
        return x, y, kx, ky
    A = rng.integers(-5, 6, size=(12, 64, 64)).astype(np.int32)
    A[3, 7] = A[3, 9]                                                        # one singular matrix
    plan = eng.plan_inverse(64, 5)
    x, y, kx, ky = both(lambda: eng.inverse_batch(A, plan=plan))
    assert ky == 0 and 1 <= kx < plan.n_primes, (kx, ky, plan.n_primes)
    assert np.array_equal(x.status, y.status) and np.array_equal(x.adj, y.adj) and np.array_equal(x.det, y.det)
    W = A.copy()
    W[0] = 5                                                                 # every entry at the declared magnitude
    x, y, kx, ky = both(lambda: eng.inverse_batch(W, plan=plan))
    assert kx == plan.n_primes
    assert np.array_equal(x.status, y.status) and np.array_equal(x.adj, y.adj) and np.array_equal(x.det, y.det)
    # rank-deficient products, solve and rref: top-(r + 1) rows
    Bm = rng.integers(-5, 6, size=(10, 40, 25)); Cm = rng.integers(-5, 6, size=(10, 25, 44))
    M = np.einsum("bik,bkj->bij", Bm, Cm).astype(np.int32)
    x, y, kx, ky = both(lambda: eng.rref_batch(M, 40))
    assert 1 <= kx
    for f in ("status", "num", "den", "pivot_col", "rank"):
        assert np.array_equal(getattr(x, f), getattr(y, f)), f
    S = M[:, :, :40].copy()
    b = rng.integers(-9, 10, size=(10, 40)).astype(np.int32)
    b[::2] = np.einsum("bij,bj->bi", S[::2], rng.integers(-2, 3, size=(5, 40)))
    pl = eng.plan_solve(40, 40, int(np.abs(S).max()), int(np.abs(b).max()), 25, 15)
    x, y, kx, ky = both(lambda: eng.solve_batch(S, b, plan=pl))
    assert 1 <= kx < pl.n_primes
    assert np.array_equal(x.status, y.status)
    ok = (x.status & 2) == 0
    assert ok.any() and not ok.all()
    for f in ("den", "particular", "generators", "pivot_col", "rank"):
        assert np.array_equal(getattr(x, f)[ok], getattr(y, f)[ok]), f
    # the fused sub-warp kernel takes the same word (thread-per-row bound kernel): config 3 shape, 4 -> 3 primes
    Bm = rng.integers(-5, 6, size=(300, 16, 10)); Cm = rng.integers(-5, 6, size=(300, 10, 16))
    S = np.einsum("bik,bkj->bij", Bm, Cm).astype(np.int32)
    b = rng.integers(-5, 6, size=(300, 16)).astype(np.int32)
    b[::2] = np.einsum("bij,bj->bi", S[::2], rng.integers(-5, 6, size=(150, 16)))
    pl = eng.plan_solve(16, 16, 250, int(np.abs(b).max()), 10, 6)
    x, y, kx, ky = both(lambda: eng.solve_batch(S, b, plan=pl))
    assert ky == 0 and 1 <= kx < pl.n_primes, (kx, pl.n_primes)
    assert np.array_equal(x.status, y.status)
    ok = (x.status & 2) == 0
    assert ok.any() and not ok.all()
    for f in ("den", "particular", "generators", "pivot_col", "rank"):
        assert np.array_equal(getattr(x, f)[ok], getattr(y, f)[ok]), f
    for mm, nn, bar in ((6, 7, 6), (12, 12, 12), (20, 24, 18), (32, 33, 32)):
        M = rng.integers(-50, 51, size=(40, mm, nn)).astype(np.int32)
        M[1] = 50                                                            # worst case for the declared magnitude
        x, y, kx, ky = both(lambda: eng.rref_batch(M, bar, a_abs_max=50, b_abs_max=50))
        for f in ("status", "num", "den", "pivot_col", "rank"):
            assert np.array_equal(getattr(x, f), getattr(y, f)), (mm, nn, f)
    # the bound is attained: det(diag) = product of the row norms; declared magnitude far above the data
    for n, v, amax in ((40, 1000, 50000), (33, 46340, 50000), (64, 3, 100)):
        D = np.zeros((3, n, n), dtype=np.int32)
        D[0] = np.diag(np.full(n, v))
        D[1] = np.diag(np.full(n, -v))
        D[2] = np.diag(rng.integers(1, v + 1, size=n))[rng.permutation(n)]
        d = eng.det_batch(D, a_abs_max=amax)
        k = eng.last_prime_count()
        got = limbs_to_ints(d.det)
        want = [v ** n, (-v) ** n, ref_port.bareiss_det(D[2].tolist())]
        assert got == want, (n, v)
        assert k == -(-(n * np.log2(v) + 1) // 30.999) or k == -(-(n * np.log2(v) + 1) // 30.999) + 1, (n, v, k)
        r = eng.inverse_batch(D, a_abs_max=amax)
        adj, det = limbs_to_ints(r.adj), limbs_to_ints(r.det)
        assert det == want
        assert [adj[0][i][i] for i in range(n)] == [v ** (n - 1)] * n


def test_in_place_inverse_kernel_equals_wide_tile_kernels(eng, monkeypatch):
    """k_tile_inv (square [A|I] in an m x m register tile: the identity column of a pivot row takes the slot of the left
    column its step eliminates) against the 2m-wide k_tile_reg and the shared-memory k_tile_elim, word for word: every
    size class incl. sizes that are no multiple of 16, forced row swaps in the first, a middle and the LAST column,
    singular inputs of several ranks (linalg.py:682-743 returns NoSolution: status bit, zero words), and the thread
    shapes behind LSX_TILE_INV_SHAPE."""
    rng = np.random.Generator(np.random.PCG64(43))
    monkeypatch.setenv("LSX_DISABLE_SUBWARP", "1")           # sizes up to 16 would take the fused sub-warp kernel
    for n in (9, 16, 17, 31, 32, 33, 47, 50, 64, 65, 96, 100, 128):
        B = 10
        lo = 5 if n <= 100 else 2                            # 128 x 128 with |a| <= 5 would need more than 32 primes
        mats = rng.integers(-lo, lo + 1, size=(B, n, n)).astype(np.int32)
        mats[1, 0, 0] = 0
        mats[1, 1, :2] = 0                                   # swaps in the first columns
        h = n // 2
        mats[2, h:, :h] = 0
        mats[2, h, h] = 0                                    # middle column: the diagonal candidate is zero
        mats[3] = np.eye(n, dtype=np.int32)[rng.permutation(n)] * rng.integers(1, lo + 1, size=(n, 1))   # swaps everywhere
        mats[4] = np.eye(n, dtype=np.int32)
        mats[4, [n - 2, n - 1]] = mats[4, [n - 1, n - 2]]    # one swap, in the second-to-last column
        mats[5, n - 1] = mats[5, 0] + mats[5, n // 3]        # rank n - 1 (entries stay within 2 lo: the plan's 32 primes)
        mats[6, h:] = 0                                      # rank n // 2
        mats[7] = np.outer(rng.choice([-1, 1], size=n), rng.choice([-1, 1], size=n))   # rank 1
        mats[8, :, n - 1] = mats[8, :, 0]                    # singular, found in the last column
        mats[9, :, 1] = 2 * mats[9, :, 0]                    # singular, found in the second column
        x = eng.inverse_batch(mats)
        monkeypatch.setenv("LSX_DISABLE_TILE_INV", "1")
        y = eng.inverse_batch(mats)
        monkeypatch.delenv("LSX_DISABLE_TILE_INV")
        monkeypatch.setenv("LSX_DISABLE_TILE_REG", "1")
        z = eng.inverse_batch(mats)
        monkeypatch.delenv("LSX_DISABLE_TILE_REG")
        for other in (y, z):
            assert np.array_equal(x.status, other.status), n
            assert np.array_equal(x.adj, other.adj) and np.array_equal(x.det, other.det), n
        assert [int(v) & 1 for v in x.status] == [0, 0, 0, 0, 0, 1, 1, 1, 1, 1], (n, x.status)
        if 32 < n <= 64:
            for shape in ("1", "2", "3"):
                monkeypatch.setenv("LSX_TILE_INV_SHAPE", shape)
                w = eng.inverse_batch(mats)
                monkeypatch.delenv("LSX_TILE_INV_SHAPE")
                assert np.array_equal(x.status, w.status) and np.array_equal(x.adj, w.adj) and np.array_equal(x.det, w.det)
    # exact values against the oracle (reference inverse = adj / det in lowest terms)
    A = rng.integers(-3, 4, size=(1, 20, 20)).astype(np.int32)
    r = eng.inverse_batch(A)
    adj, det = limbs_to_ints(r.adj), limbs_to_ints(r.det)
    want = ref_port.inverse(A[0].tolist())
    assert want is not None and [[Fraction(v, det[0]) for v in row] for row in adj[0]] == want
