"""Several GPUs behind one context of the C-ABI (lsx_create_multi): by-matrix slicing of host batches and the
by-prime determinant with the NCCL all-gather inside the library.  Needs two GPUs (`gpurun --gpus 2`); on one GPU
only the single-device form of lsx_det_large is exercised."""
import numpy as np
import pytest

from linalg_solver_b200.convert import limbs_to_ints
from oracle import golden_io

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_det_large_single_device_context():
    from linalg_solver_b200 import Engine
    eng = Engine(0)
    try:
        for c in golden_io.load("c5_standins")["cases"]:
            if c["n"] > 256:
                continue
            rng = np.random.Generator(np.random.PCG64(c["seed"]))
            A = rng.integers(-5, 6, size=(c["n"], c["n"]), dtype=np.int64)
            det, primes = eng.det_large(A)
            assert det == int(c["det"]) and primes >= 1
    finally:
        eng.close()


@pytest.mark.skipif("_n_gpus() < 2")
def test_multi_context_matches_single_device():
    from linalg_solver_b200 import Engine
    one, two = Engine(0), Engine(devices=[0, 1])
    try:
        rng = np.random.Generator(np.random.PCG64(12))
        # by matrix: odd batch (uneven slices), every operation, host arrays
        A8 = rng.integers(-5, 6, size=(100001, 8, 8), dtype=np.int32)
        A8[7, 3] = A8[7, 1]
        r1, r2 = one.inverse_batch(A8, a_abs_max=5), two.inverse_batch(A8, a_abs_max=5)
        assert np.array_equal(r1.adj, r2.adj) and np.array_equal(r1.det, r2.det) and np.array_equal(r1.status, r2.status)
        i8 = A8.astype(np.int8)
        r3 = two.inverse_batch(i8, a_abs_max=5)
        assert np.array_equal(r1.adj, r3.adj) and np.array_equal(r1.status, r3.status)
        Bm = rng.integers(-5, 6, size=(4097, 16, 10), dtype=np.int64)
        Cm = rng.integers(-5, 6, size=(4097, 10, 16), dtype=np.int64)
        A16 = np.einsum("bik,bkj->bij", Bm, Cm).astype(np.int32)
        b16 = rng.integers(-5, 6, size=(4097, 16), dtype=np.int32)
        s1 = one.solve_batch(A16, b16, max_rank=10, gen_cap=6)
        s2 = two.solve_batch(A16, b16, max_rank=10, gen_cap=6)
        for f in ("den", "particular", "generators", "pivot_col", "rank", "status"):
            assert np.array_equal(getattr(s1, f), getattr(s2, f)), f
        A4 = rng.integers(-5, 6, size=(1, 4, 4), dtype=np.int32)       # fewer matrices than GPUs
        assert np.array_equal(one.det_batch(A4).det, two.det_batch(A4).det)
        q1, q2 = one.rref_batch(A16[:33], 16), two.rref_batch(A16[:33], 16)
        assert np.array_equal(q1.num, q2.num) and np.array_equal(q1.pivot_col, q2.pivot_col)
        assert np.array_equal(one.rank_batch(A16[:33]).rank, two.rank_batch(A16[:33]).rank)
        # device tensors belong to one GPU
        import torch
        from linalg_solver_b200 import LsxError
        with pytest.raises(LsxError):
            two.inverse_batch(torch.from_numpy(A8[:16]).cuda(), a_abs_max=5)
        # by prime: the DomainMatrix stand-ins (tile kernel per prime at 128, blocked tensor-core LU at 256 / 512)
        for c in golden_io.load("c5_standins")["cases"]:
            if c["n"] not in (128, 256, 512):
                continue
            A = np.random.Generator(np.random.PCG64(c["seed"])).integers(-5, 6, size=(c["n"], c["n"]), dtype=np.int64)
            det, primes = two.det_large(A)
            assert det == int(c["det"]), c["n"]
        assert two.launch_count > 0
        from linalg_solver_b200._lib import lib
        assert lib.lsx_device_count(two._ctx) == 2 and lib.lsx_multi_uses_nccl(two._ctx) == 1   # NCCL all-gather ran
    finally:
        one.close()
        two.close()
