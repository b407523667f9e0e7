"""Sharding of independent source blocks over ranks (one process per GPU).

A RaptorQ object is cut into Z source blocks that never interact (reference
lib/nanorq.c:57,130-146: one block_encoder per SBN), so the multi-GPU path is a
partition with NO data-path collective: source block `sbn` belongs to rank
`sbn % world`.  The only cross-rank traffic is the timing reduction of a
benchmark (max over ranks) and, for tests, gathering per-block digests."""
import torch
import torch.distributed as dist


def blocks_for_rank(n_blocks, rank, world):
    """SBNs owned by `rank`: sbn % world == rank (SBN b -> device b mod n)."""
    if not (0 <= rank < world):
        raise ValueError("rank %r outside world %r" % (rank, world))
    return list(range(rank, n_blocks, world))


def owner_of(sbn, world):
    return sbn % world


def max_over_ranks(x, device="cpu"):
    """Slowest rank's value (the time a sharded job really takes)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_objects(obj):
    """[obj of rank 0, obj of rank 1, ...] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
