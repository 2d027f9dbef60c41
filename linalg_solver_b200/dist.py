"""Multi-GPU layout of the elimination path: one process per GPU, no data-path collective for
batches (independent matrices), one all-gather of residues for a single large determinant.

SURVEY.md section 8e: batches are sharded BY MATRIX in contiguous slices, every rank runs all
primes and the CRT for its slice and the results stay sharded; a single large matrix is sharded
BY PRIME (each rank holds its own int32 copy of A), the per-prime residues are all-gathered
(NCCL over NVLink on GPUs, gloo in the CPU tests) and the CRT runs on every rank.
"""
from typing import List, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [begin, end) of `total` units owned by `rank` (sizes differ by at most 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def all_gather_residues(local, total: int, group=None):
    """All-gather the per-prime residues of a by-prime sharded determinant.

    `local` is this rank's 1-D tensor for primes shard_range(total, rank, world) (int32 storage of
    uint32 residues, on the GPU under NCCL, on the CPU under gloo).  Returns the full tensor of
    `total` residues in prime order on every rank.  Uneven shards are padded to the largest shard
    so that one fixed-size all_gather suffices (about 4 B x 1100 primes for config 5).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    width = max(sizes) if sizes else 0
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = torch.empty(world * width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = [out[r * width: r * width + sizes[r]] for r in range(world)]
    return torch.cat(parts) if parts else out


def _gather_any(local, total: int, device_index, group=None):
    """all_gather_residues for a residue vector in either container: torch tensors go through as they are; a
    numpy array (host-memory engine call) is wrapped as an int32 tensor -- on the engine's GPU when the group's
    backend is NCCL, which only moves CUDA tensors -- and comes back as a numpy uint32 array."""
    import numpy as np
    import torch
    import torch.distributed as dist

    if not isinstance(local, np.ndarray):
        return all_gather_residues(local, total, group)
    t = torch.from_numpy(np.ascontiguousarray(local).view(np.int32).copy())
    if dist.get_backend(group) == "nccl":
        t = t.to(torch.device("cuda", device_index))
    full = all_gather_residues(t, total, group)
    return full.cpu().numpy().view(np.uint32)


def det_large_sharded(engine, A, a_abs_max: int = None, group=None, sharded: bool = True):
    """Determinant of one large integer matrix, primes sharded over the ranks of `group`.

    Returns (det_words, n_primes): `det_words` is the signed determinant as little-endian 32-bit
    words (two's complement) on every rank.  The prime count comes from the worst-case Hadamard bound of
    `a_abs_max` when it is given, else from the row/column norms of A itself (every rank holds the same A,
    so every rank derives the same count).

    With `sharded` (the default) and an initialised process group this is a COLLECTIVE: every rank of `group`
    must call it with the same matrix.  `sharded=False` computes all primes on this rank and never communicates
    (what ``Matrix.determinant`` does: a method of one object must not turn into an implicit collective).
    A may be a torch tensor (CUDA: device call) or a numpy array (host call); the result follows the input.
    """
    import torch.distributed as dist

    n = A.shape[0]
    if a_abs_max is None:
        K, bits = engine.det_large_prime_count_for(A)
    else:
        K, bits = engine.det_large_prime_count(n, a_abs_max)
    if sharded and dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    b, e = shard_range(K, rank, world)
    local = engine.det_large_residues(A, b, e - b)
    full = _gather_any(local, K, getattr(engine, "device", 0), group) if world > 1 else local
    limbs = int(bits + 2) // 32 + 1
    return engine.crt_signed(full, limbs), K
