"""Out-of-bounds writes: every output buffer of every batched operation sits between guard words that must come back
untouched, for each kernel family (fused 8x8, sub-warp, shared-memory tile, register tile, lowest terms).
compute-sanitizer is closed on this pool (profiles/r02d_sanitizer_closed.txt); this is the check that remains.
Also covers caller pointers that are not 16-byte aligned (a device view with an odd element offset is legal)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD = 64            # words on each side
SENT = 0x5A5A5A5A


@pytest.fixture(scope="module")
def eng():
    from linalg_solver_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


class Guarded:
    """Device int32 buffer of `shape` between two guard zones, optionally shifted by `skew` words (misalignment)."""

    def __init__(self, shape, skew=0):
        import torch
        self.n = int(np.prod(shape))
        self.skew = skew
        self.raw = torch.full((GUARD + skew + self.n + GUARD,), SENT, dtype=torch.int32, device="cuda")
        self.view = self.raw[GUARD + skew: GUARD + skew + self.n].view(*shape)

    def intact(self):
        lo = self.raw[: GUARD + self.skew]
        hi = self.raw[GUARD + self.skew + self.n:]
        return bool((lo == SENT).all()) and bool((hi == SENT).all())


def _out_like(res_cls, plan, fields, skew=0):
    bufs = {k: Guarded(shape, skew) for k, shape in fields.items()}
    kw = {k: b.view for k, b in bufs.items()}
    return res_cls(plan=plan, **kw), bufs


@pytest.mark.parametrize("n,batch,amax", [(8, 1000, 5), (8, 129, 5), (5, 77, 5), (3, 33, 9), (1, 5, 7), (8, 300, 3000), (12, 50, 5),
                                          (64, 6, 5)])
@pytest.mark.parametrize("skew", [0, 1])
def test_inverse_outputs_stay_inside_their_buffers(eng, n, batch, amax, skew):
    import torch
    from linalg_solver_b200.engine import InverseResult
    rng = np.random.Generator(np.random.PCG64(n * 1000 + batch))
    A_host = rng.integers(-amax, amax + 1, size=(batch, n, n), dtype=np.int32)
    A_host[batch // 2, -1] = A_host[batch // 2, 0]                       # a singular matrix
    ref = eng.inverse_batch(A_host, a_abs_max=amax)
    plan = ref.plan
    L = plan.limbs
    A = Guarded((batch, n, n), skew)
    A.view.copy_(torch.from_numpy(A_host))
    out, bufs = _out_like(InverseResult, plan, {"adj": (batch, n, n, L), "det": (batch, L), "status": (batch,)}, skew)
    eng.inverse_batch(A.view, plan=plan, out=out)
    torch.cuda.synchronize()
    assert all(b.intact() for b in bufs.values()) and A.intact()
    assert np.array_equal(out.adj.cpu().numpy().view(np.uint32), ref.adj)
    assert np.array_equal(out.det.cpu().numpy().view(np.uint32), ref.det)
    assert np.array_equal(out.status.cpu().numpy(), ref.status)


@pytest.mark.parametrize("m,n,batch,rank", [(4, 4, 999, 4), (16, 16, 257, 10), (7, 9, 100, 5), (20, 30, 40, 20), (40, 41, 9, 33), (64, 64, 5, 48)])
def test_solve_rref_det_rank_outputs_stay_inside_their_buffers(eng, m, n, batch, rank):
    import torch
    from linalg_solver_b200.engine import DetResult, RankResult, RrefResult, SolveResult
    rng = np.random.Generator(np.random.PCG64(m * 100 + n))
    Bm = rng.integers(-3, 4, size=(batch, m, rank), dtype=np.int64)
    Cm = rng.integers(-3, 4, size=(batch, rank, n), dtype=np.int64)
    A_host = np.einsum("bik,bkj->bij", Bm, Cm).astype(np.int32)
    b_host = rng.integers(-3, 4, size=(batch, m), dtype=np.int32)
    amax = int(np.abs(A_host).max())
    # find_preimage_of
    ref = eng.solve_batch(A_host, b_host, a_abs_max=amax, b_abs_max=3, gen_cap=min(n, 8))
    plan, L, G = ref.plan, ref.plan.limbs, ref.plan.gen_cap
    A, b = Guarded((batch, m, n)), Guarded((batch, m))
    A.view.copy_(torch.from_numpy(A_host))
    b.view.copy_(torch.from_numpy(b_host))
    out, bufs = _out_like(SolveResult, plan, {"den": (batch, L), "particular": (batch, n, L), "generators": (batch, n, G, L),
                                             "pivot_col": (batch, plan.pivot_slots), "rank": (batch,), "status": (batch,)})
    eng.solve_batch(A.view, b.view, plan=plan, out=out)
    torch.cuda.synchronize()
    assert all(x.intact() for x in bufs.values()) and A.intact() and b.intact()
    for f in ("den", "particular", "generators", "pivot_col", "rank", "status"):
        got = getattr(out, f).cpu().numpy()
        want = getattr(ref, f)
        assert np.array_equal(got.view(want.dtype), want), f
    # row_reduce
    ref = eng.rref_batch(A_host, n, a_abs_max=amax)
    plan, L = ref.plan, ref.plan.limbs
    out, bufs = _out_like(RrefResult, plan, {"num": (batch, m, n, L), "den": (batch, L), "pivot_col": (batch, plan.pivot_slots),
                                            "rank": (batch,), "status": (batch,)})
    eng.rref_batch(A.view, n, plan=plan, out=out)
    torch.cuda.synchronize()
    assert all(x.intact() for x in bufs.values()) and A.intact()
    assert np.array_equal(out.num.cpu().numpy().view(np.uint32), ref.num) and np.array_equal(out.rank.cpu().numpy(), ref.rank)
    # rank, and det for square shapes
    ref = eng.rank_batch(A_host, a_abs_max=amax)
    out, bufs = _out_like(RankResult, ref.plan, {"rank": (batch,), "status": (batch,)})
    eng.rank_batch(A.view, plan=ref.plan, out=out)
    torch.cuda.synchronize()
    assert all(x.intact() for x in bufs.values()) and np.array_equal(out.rank.cpu().numpy(), ref.rank)
    if m == n:
        ref = eng.det_batch(A_host, a_abs_max=amax)
        out, bufs = _out_like(DetResult, ref.plan, {"det": (batch, ref.plan.limbs), "rank": (batch,), "status": (batch,)})
        eng.det_batch(A.view, plan=ref.plan, out=out)
        torch.cuda.synchronize()
        assert all(x.intact() for x in bufs.values()) and np.array_equal(out.det.cpu().numpy().view(np.uint32), ref.det)


def test_lowest_terms_outputs_stay_inside_their_buffers(eng):
    import torch
    rng = np.random.Generator(np.random.PCG64(5))
    for L, B, C in [(1, 100, 7), (3, 33, 5), (11, 9, 64), (20, 4, 3)]:
        num = Guarded((B, C, L))
        den = Guarded((B, L))
        num.view.copy_(torch.from_numpy(rng.integers(-2**31, 2**31, size=(B, C, L), dtype=np.int64).astype(np.int32)))
        den.view.copy_(torch.from_numpy(rng.integers(-2**31, 2**31, size=(B, L), dtype=np.int64).astype(np.int32)))
        # results are allocated by the call; run it on guarded inputs and check the inputs' guards and idempotence
        p, q = eng.lowest_terms(num.view, den.view)
        torch.cuda.synchronize()
        assert num.intact() and den.intact()
        p2, q2 = eng.lowest_terms(p.reshape(B * C, 1, L), q.reshape(B * C, L))
        assert torch.equal(p2.reshape(p.shape), p) and torch.equal(q2.reshape(q.shape), q)
