"""Batch sharding across the GPUs of one box.

Witnesses are independent (the reference itself is one process per witness, common/main.cpp:334-371), so
the batch is split contiguously by rank, the program and the CSR matrix are replicated, and NO collective is
needed while computing.  The only exchange is the final gather of the per-witness flags (status word, first
violated constraint) to rank 0.  Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU
tests)."""
from __future__ import annotations


def shard_range(total, rank, world):
    """Contiguous [begin, end) of `total` witnesses owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_flags(local_flags, total, group=None):
    """local_flags: 1-D int32 tensor holding this rank's shard (in shard_range order).
    Returns the full [total] tensor on rank 0 and None elsewhere."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return local_flags
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    assert local_flags.numel() == sizes[rank], "shard size mismatch"
    cap = max(sizes) if sizes else 0
    padded = torch.zeros(cap, dtype=local_flags.dtype, device=local_flags.device)
    padded[:sizes[rank]] = local_flags
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)          # tiny: 4 bytes per witness
    if rank != 0:
        return None
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)])


def count_failures(local_status, local_first_bad, group=None):
    """Sum over ranks of witnesses with a non-zero status or a violated constraint (one all-reduce of 2 ints)."""
    import torch
    import torch.distributed as dist

    t = torch.stack([(local_status != 0).sum(), (local_first_bad != -1).sum()]).to(torch.int64)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, group=group)
    return int(t[0]), int(t[1])
