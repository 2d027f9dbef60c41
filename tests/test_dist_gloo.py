"""Host-side multi-rank logic on CPU: world_size-2 gloo run of the by-prime residue all-gather and
the by-matrix shard arithmetic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from linalg_solver_b200.dist import all_gather_residues, shard_range, shard_sizes


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 8, 1100, 2**20):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = shard_range(total, rank, world)
        local = torch.arange(b, e, dtype=torch.int32) * 3 + 1
        full = all_gather_residues(local, total)
        q.put((rank, full.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [7, 1100])
def test_all_gather_residues_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [i * 3 + 1 for i in range(total)]
    for _, full in got:
        assert full == want


class _HostEngine:
    """Stand-in for linalg_solver_b200.Engine on a box without a GPU: the same three calls det_large_sharded makes,
    answered by the CPU oracle (numpy containers, like Engine's host-memory calls)."""
    device = 0

    def __init__(self, primes):
        self.p = primes

    def det_large_prime_count_for(self, A):
        return len(self.p), 31.0 * len(self.p) - 2.0

    def det_large_residues(self, A, begin, count):
        import numpy as np
        from oracle.det_mod_p import det_mod_p
        return np.array([det_mod_p(A, int(q)) for q in self.p[begin:begin + count]], dtype=np.uint32)

    def crt_signed(self, residues, limbs):
        from linalg_solver_b200.convert import crt_basis
        M, coef = crt_basis([int(q) for q in self.p])
        x = sum(int(r) * c for r, c in zip(residues, coef)) % M
        return x - M if x > M // 2 else x


def _worker_det(rank, world, port, q):
    import numpy as np
    from linalg_solver_b200.dist import det_large_sharded
    from tests.device_model import prime_table
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A = np.random.Generator(np.random.PCG64(5)).integers(-5, 6, size=(12, 12), dtype=np.int32)
        eng = _HostEngine(prime_table(3))
        det, K = det_large_sharded(eng, A)                       # numpy input, sharded over the two ranks
        alone, _ = det_large_sharded(eng, A, sharded=False)      # no communication
        q.put((rank, int(det), int(alone), K))
    finally:
        dist.destroy_process_group()


def test_det_large_sharded_numpy_input_gloo_world2():
    """the by-prime route with a numpy matrix under an initialised process group."""
    import numpy as np
    from oracle import ref_port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_det, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A = np.random.Generator(np.random.PCG64(5)).integers(-5, 6, size=(12, 12), dtype=np.int32)
    want = ref_port.bareiss_det(A.tolist())
    for _, det, alone, K in got:
        assert det == want and alone == want and K == 3
