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


# ---- FASTQ text (the FASTQ-in -> trimmed-FASTQ-out paths of atropos_b200/fastq.py) --------------------------------
def _line_start(buf, pos):
    """first line start at or after pos (pos itself if it follows a newline or is 0)"""
    n = len(buf)
    if pos <= 0:
        return 0
    if pos >= n:
        return n
    if buf[pos - 1] == 10:
        return pos
    nxt = bytes(buf[pos:min(n, pos + (1 << 20))]).find(b"\n")
    if nxt < 0:
        nxt = bytes(buf[pos:]).find(b"\n")
        if nxt < 0:
            return n
    return pos + nxt + 1


def _next_lines(buf, pos, count):
    """start offsets of the `count` lines beginning at line start `pos` (shorter at the end of the text)"""
    n = len(buf)
    out = []
    while len(out) < count and pos < n:
        out.append(pos)
        nxt = bytes(buf[pos:min(n, pos + (1 << 16))]).find(b"\n")
        if nxt < 0:
            nxt = bytes(buf[pos:]).find(b"\n")
            if nxt < 0:
                break
        pos += nxt + 1
    return out


def fastq_record_start(buf, pos):
    """The first record start at or after byte `pos` of a well-formed 4-line FASTQ text (uint8 array / bytes).

    A line starts a record iff it begins with '@' and the line two below begins with '+'; a quality line that
    happens to begin with '@' fails that test because the line two below it is a sequence. (Sequence lines are
    assumed not to begin with '+': true for nucleotide / IUPAC data; the text as a whole is still validated by the
    GPU reader, so a wrong guess can only turn into a FormatError, never into silently different output.)"""
    n = len(buf)
    p = _line_start(buf, pos)
    for _ in range(8):
        if p >= n:
            return n
        ls = _next_lines(buf, p, 3)
        if len(ls) == 3 and buf[ls[0]] == 64 and buf[ls[2]] == 43:
            return p
        if len(ls) < 2:
            return n
        p = ls[1]
    raise ValueError("no FASTQ record start found near byte %d" % pos)


def fastq_split_points(buf, parts):
    """parts+1 byte offsets cutting the text into `parts` runs of whole records of about equal size."""
    n = len(buf)
    pts = [0]
    for r in range(1, parts):
        pts.append(max(pts[-1], fastq_record_start(buf, n * r // parts)))
    pts.append(n)
    return pts


def fastq_pair_split_points(buf1, buf2, parts):
    """Split two paired texts at the same RECORD numbers: file 1 is cut near equal byte counts, file 2 at the record
    with the same index, found by counting newlines (both texts hold 4 lines per record)."""
    import numpy as np
    p1 = fastq_split_points(buf1, parts)
    a1, a2 = np.frombuffer(buf1, dtype=np.uint8), np.frombuffer(buf2, dtype=np.uint8)
    nl2 = np.flatnonzero(a2 == 10)
    p2 = [0]
    for cut in p1[1:-1]:
        lines = int(np.count_nonzero(a1[:cut] == 10))            # newlines before the cut = 4 * records
        p2.append(int(nl2[lines - 1]) + 1 if 0 < lines <= len(nl2) else (0 if lines == 0 else len(a2)))
    p2.append(len(a2))
    return p1, p2


def gather_trim_stats(stats):
    """Sum a fastq.TrimStats / fastq.PairTrimStats over all ranks (the reference merges its workers' summaries the
    same way, commands/multicore.py:368-389). No-op without an initialised process group."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return stats
    nccl = dist.get_backend() == "nccl"

    def allsum(x):
        t = torch.as_tensor(np.ascontiguousarray(x)).clone()
        if nccl:
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    for name in ("errors_front", "errors_back", "adjacent"):
        if hasattr(stats, name):
            v = getattr(stats, name)
            if isinstance(v, list):
                for i in range(len(v)):
                    v[i][...] = allsum(v[i])
            else:
                v[...] = allsum(v)
    scalars = [k for k in ("records", "with_adapters", "bp_in", "bp_out", "overflow", "insert_matches") if hasattr(stats, k)]
    flat = []
    for k in scalars:
        v = getattr(stats, k)
        flat.extend(v if isinstance(v, list) else [v])
    ops_keys = sorted(stats.ops)
    for k in ops_keys:
        v = stats.ops[k]
        flat.extend(v if isinstance(v, list) else [v])
    tot = [int(x) for x in allsum(np.array(flat, dtype=np.int64))]
    i = 0
    for k in scalars:
        v = getattr(stats, k)
        if isinstance(v, list):
            setattr(stats, k, tot[i:i + len(v)]); i += len(v)
        else:
            setattr(stats, k, tot[i]); i += 1
    for k in ops_keys:
        v = stats.ops[k]
        if isinstance(v, list):
            stats.ops[k] = tot[i:i + len(v)]; i += len(v)
        else:
            stats.ops[k] = tot[i]; i += 1
    return stats
