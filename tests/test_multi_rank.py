"""N>1 path on CPU: two gloo ranks shard a batch, each produces its flags, rank 0 gathers them."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from circom_cvm_b200.sharding import count_failures, gather_flags, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                b, e = shard_range(total, r, world)
                covered += list(range(b, e)) if total <= 1000 else []
                assert 0 <= b <= e <= total
            sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
            if total <= 1000:
                assert covered == list(range(total))
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_range(total, rank, world)
    idx = torch.arange(b, e, dtype=torch.int32)
    status = (idx % 97 == 0).to(torch.int32)                  # "assert failed" on every 97th witness
    first_bad = torch.where(idx % 101 == 0, idx, torch.full_like(idx, -1))
    full = gather_flags(status, total)
    fails = count_failures(status, first_bad)
    if rank == 0:
        torch.save({"full": full, "fails": fails}, out)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_of_flags(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    total = 1001
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, total, out), nprocs=2, join=True)
    got = torch.load(out)
    idx = torch.arange(total, dtype=torch.int32)
    assert torch.equal(got["full"], (idx % 97 == 0).to(torch.int32))
    assert got["fails"] == (int((idx % 97 == 0).sum()), int((idx % 101 == 0).sum()))
