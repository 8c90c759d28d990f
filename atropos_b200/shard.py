"""Multi-GPU sharding of the read index range: reads are independent, so rank r of
`world` aligns the contiguous range shard_range(n, world, r) on its own GPU and no data-path collective is
needed. The reference does the same with 1000-read batches over worker processes
(atropos/commands/multicore.py:164-232) and merges per-worker summaries on the host (:368-389).

`gather_counts` is the only exchange: a host-side merge of a few per-shard counters (torch.distributed,
any backend) -- the analogue of the reference's Summary.merge.
"""


def shard_range(n, world, rank):
    """Contiguous, balanced [start, stop): the first n % world shards get one extra read."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(int(n), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(ascii, offsets, world, rank):
    """Slice an (ascii, offsets) batch to this rank's shard; offsets are re-based to start at 0."""
    n = len(offsets) - 1
    s, e = shard_range(n, world, rank)
    lo, hi = int(offsets[s]), int(offsets[e])
    return ascii[lo:hi], offsets[s:e + 1] - offsets[s], (s, e)


def summarize(records):
    """Per-shard counters from a MATCH_DTYPE array: (reads, reads_with_adapter, total_errors, bases_removed_back)."""
    import numpy as np
    hit = records["status"] == 1
    return np.array([len(records), int(hit.sum()), int(records["errors"][hit].sum())], dtype=np.int64)


def gather_counts(local_counts):
    """Sum per-shard counters over all ranks (no-op without an initialised process group)."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(local_counts).clone()
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t = t.cpu()
    return t.numpy()
