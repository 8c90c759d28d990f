"""N > 1 host logic on CPU: world_size-2 gloo process group. Each rank takes its shard of one batch, "aligns"
it (here with the CPU oracle standing in for the GPU, since this container has none -- the sharding logic is
what is under test), and the concatenation of the shards must equal the unsharded result; counters merge by
all-reduce like the reference's Summary.merge."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC"


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from atropos_b200 import shard, synth
    from oracle import oracle
    n, L = 5001, 100
    reads = synth.synth_se(n, L, T1, seed=7, device="cpu").numpy().reshape(-1)
    offsets = np.arange(n + 1, dtype=np.int64) * L
    a, o, (s, e) = shard.shard_batch(reads, offsets, world, rank)
    res = oracle.locate_batch(T1, a, o, 0.1, oracle.BACK, False, False, 3, 1)
    np.save(os.path.join(tmpdir, "shard%d.npy" % rank), res)
    counts = np.array([e - s, int((res[:, 0] == 1).sum()), int(res[res[:, 0] == 1, 6].sum())], dtype=np.int64)
    total = shard.gather_counts(counts)
    if rank == 0:
        np.save(os.path.join(tmpdir, "total.npy"), total)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partition():
    from atropos_b200 import shard
    for n in (0, 1, 7, 100, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard.shard_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_two_rank_gloo_sharding(tmp_path):
    world, port = 2, 29641
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from atropos_b200 import synth
    from oracle import oracle
    n, L = 5001, 100
    reads = synth.synth_se(n, L, T1, seed=7, device="cpu").numpy().reshape(-1)
    offsets = np.arange(n + 1, dtype=np.int64) * L
    full = oracle.locate_batch(T1, reads, offsets, 0.1, oracle.BACK, False, False, 3, 1)
    parts = np.concatenate([np.load(tmp_path / ("shard%d.npy" % r)) for r in range(world)])
    assert np.array_equal(parts, full)
    total = np.load(tmp_path / "total.npy")
    hit = full[:, 0] == 1
    assert list(total) == [n, int(hit.sum()), int(full[hit, 6].sum())]
