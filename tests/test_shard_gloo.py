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


# ---- FASTQ text sharding -----------------------------------------------------------------------------------------
def test_fastq_split_points_are_record_starts():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fastq_cases
    from atropos_b200 import shard
    for case in fastq_cases.cases():
        if "error" in case["result"] or "\r" in case["text"]:
            continue
        text = case["text"].encode("latin-1")
        starts = set()
        pos = 0
        lines = text.split(b"\n")
        for i, ln in enumerate(lines):
            if i % 4 == 0:
                starts.add(pos)
            pos += len(ln) + 1
        starts.add(len(text))
        for parts in (1, 2, 3, 8):
            pts = shard.fastq_split_points(np.frombuffer(text, dtype=np.uint8), parts)
            assert pts[0] == 0 and pts[-1] == len(text) and all(a <= b for a, b in zip(pts, pts[1:]))
            assert all(p in starts for p in pts), (case["label"], parts)
    # a quality line that starts with '@' right at the probe position must not be taken for a header
    rec = b"@r\nACGT\n+\n@III\n"
    text = rec * 50
    for pos in range(0, len(text), 3):
        assert shard.fastq_record_start(np.frombuffer(text, dtype=np.uint8), pos) % len(rec) == 0


def _fastq_worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pickle
    import fastq_cases
    import hostsim
    from atropos_b200 import shard
    # single-end: every rank trims its run of whole records (the CPU simulation stands in for the GPU here)
    case = [c for c in fastq_cases.cases() if c["label"] == "ops_panel_times2_all"][0]
    adapters = fastq_cases.adapters_of(case)
    text = case["text"].encode("latin-1")
    pts = shard.fastq_split_points(np.frombuffer(text, dtype=np.uint8), world)
    out, stats, consumed = hostsim.trim_fastq(text[pts[rank]:pts[rank + 1]], adapters, times=case["times"], **case["read_ops"])
    assert consumed == pts[rank + 1] - pts[rank]
    stats = shard.gather_trim_stats(stats)
    # paired-end: both files cut at the same record numbers
    pcase = [c for c in fastq_cases.pe_cases() if c["label"] == "ops_quality_cut_maxn"][0]
    a1, a2, ia = fastq_cases.pe_objects(pcase)
    t1, t2 = pcase["text1"].encode("latin-1"), pcase["text2"].encode("latin-1")
    p1, p2 = shard.fastq_pair_split_points(t1, t2, world)
    pouts, pstats, _ = hostsim.trim_fastq_pe(t1[p1[rank]:p1[rank + 1]], t2[p2[rank]:p2[rank + 1]], a1, a2, ia, **pcase["read_ops"])
    pstats = shard.gather_trim_stats(pstats)
    with open(os.path.join(tmpdir, "fq%d.pkl" % rank), "wb") as fh:
        pickle.dump({"out": out, "stats": stats if rank == 0 else None, "pout": pouts, "pstats": pstats if rank == 0 else None}, fh)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_fastq_sharding(tmp_path):
    import pickle
    world, port = 2, 29653
    mp.spawn(_fastq_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fastq_cases
    parts = [pickle.load(open(tmp_path / ("fq%d.pkl" % r), "rb")) for r in range(world)]
    case = [c for c in fastq_cases.cases() if c["label"] == "ops_panel_times2_all"][0]
    fastq_cases.check(case, b"".join(p["out"] for p in parts), parts[0]["stats"], fastq_cases.adapters_of(case))
    pcase = [c for c in fastq_cases.pe_cases() if c["label"] == "ops_quality_cut_maxn"][0]
    fastq_cases.pe_check(pcase, (b"".join(p["pout"][0] for p in parts), b"".join(p["pout"][1] for p in parts)), parts[0]["pstats"])
