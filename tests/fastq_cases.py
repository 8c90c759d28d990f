"""Shared by the CPU (host simulator) and GPU tests of the FASTQ path: golden cases produced by the reference
command line (tests/golden/make_fastq_golden.py) and the comparison of a result with them."""
import golden_util
from atropos_b200.adapters import ANYWHERE, BACK, FRONT, PREFIX, Adapter, LinkedAdapter

WHERE = {"back": BACK, "front": FRONT, "anywhere": ANYWHERE}


def cases():
    return golden_util.load("fastq_trim")


def adapters_of(case):
    """Adapter objects in the order the reference's AdapterCutter tried them (the order of its report)."""
    res = case["result"]
    if any(w == "linked" for _, w in case["adapters"]):          # "-a FRONT...BACK": one LinkedAdapter
        front, back = case["adapters"][0][0].split("...")
        return [LinkedAdapter(front, back, max_error_rate=case["error_rate"], min_overlap=case["overlap"])]
    if "adapters" in res:
        order = [(a["sequence"], a["where"]) for a in res["adapters"]]
    else:
        order = [tuple(a) for a in case["adapters"]]
    return [Adapter(seq, WHERE[w], max_error_rate=case["error_rate"], min_overlap=case["overlap"]) for seq, w in order]


def _str_keys(d):
    return {str(k): ({str(k2): v2 for k2, v2 in v.items()} if isinstance(v, dict) else v) for k, v in d.items()}


def check_ops(case, stats, paired=False):
    """Trimmer.trimmed_bases / FilterWrapper.filtered / records written, for every modifier and filter that was on"""
    gold = case["result"].get("ops")
    if gold is None:
        return
    for key, val in gold.items():
        if val is None:                                   # that modifier / filter was not part of the command
            assert stats.ops[key] in (0, [0, 0]), (case["label"], key)
        elif isinstance(val, list):
            assert stats.ops[key][:len(val)] == val and not any(stats.ops[key][len(val):]), (case["label"], key)
        else:
            assert stats.ops[key] == val, (case["label"], key)


def check(case, out, stats, adapters):
    res = case["result"]
    check_ops(case, stats)
    assert bytes(out) == res["out"].encode("latin-1")
    assert stats.records == res["records"]
    assert stats.with_adapters == res["with_adapters"]
    assert stats.bp_in == res["bp_in"]
    if res["bp_out"] is not None:
        assert stats.bp_out == res["bp_out"]
    assert stats.overflow == 0
    if len(adapters) == 1 and isinstance(adapters[0], LinkedAdapter):
        for a, gold in enumerate(res["adapters"]):               # index 0 = the front adapter, 1 = the back adapter
            mine = stats.adapter_summary(a, ANYWHERE)             # both histograms of that sub-adapter
            for key in ("lengths_front", "lengths_back", "errors_front", "errors_back"):
                assert _str_keys(mine[key]) == gold[key], (case["label"], a, key)
        return
    for a, (ad, gold) in enumerate(zip(adapters, res["adapters"])):
        mine = stats.adapter_summary(a, ad.where)
        for key in ("lengths_front", "lengths_back", "errors_front", "errors_back", "adjacent_bases"):
            assert (key in gold) == (key in mine), (case["label"], a, key)
            if key in gold:
                assert _str_keys(mine[key]) == gold[key], (case["label"], a, key)


# ---- paired-end ("--aligner insert") -----------------------------------------------------------------------------
def pe_cases():
    return golden_util.load("fastq_trim_pe")


def pe_objects(case):
    """(adapter1, adapter2, insert_aligner) as the reference's command line builds them in insert mode
    (trim/cli.py:667-684, :801-802; trim/__init__.py:356-371, 444-456)."""
    from atropos_b200 import synth
    from atropos_b200.align import InsertAligner
    from atropos_b200.util import RandomMatchProbability
    if case.get("mode") == "adapter":
        # --aligner adapter: indel cost 1, overlap 3, no max_rmp (trim/cli.py:659-666); adapters in the report's order
        lists = [[Adapter(a["sequence"], WHERE[a["where"]], max_error_rate=case["error_rate"], min_overlap=3) for a in per_read]
                 for per_read in case["result"]["adapters"]]
        return (lists[0] or None), (lists[1] or None), None
    e = case["error_rate"]
    insert_rate = e or 0.2            # evaluated before error_rate gets its 0.1 default
    adapter_rate = 0.1 if e is None else e
    rmp = RandomMatchProbability()
    kw = dict(max_error_rate=adapter_rate, min_overlap=1, indel_cost=3, max_rmp=1e-6, match_probability=rmp)
    a1 = Adapter(synth.TRUSEQ_R1, BACK, **kw)
    a2 = Adapter(synth.TRUSEQ_R2, BACK, **kw)
    ia = InsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, match_probability=rmp, max_insert_mismatch_frac=insert_rate,
                       max_adapter_mismatch_frac=insert_rate)
    return a1, a2, ia


def merge_kwargs(case):
    """FastqPairTrimmer / hostsim keyword arguments of a case run with --merge-overlapping (trim/cli.py:686-691: the merge
    error rate defaults to -e, else 0.2; --merge-min-overlap to 0.9)"""
    m = case.get("merge")
    if m is None:
        return {}
    return dict(merge_overlapping=True, merge_min_overlap=m.get("min_overlap", 0.9),
                merge_error_rate=m.get("error_rate", case["error_rate"] or 0.2), merged_output=m.get("output", True))


def pe_check(case, outs, stats):
    res = case["result"]
    merged = res.get("merged")
    if merged:
        # Formatters.summarize (writers.py:161-170) counts the merged output's records and bases with the paired outputs
        assert bytes(outs[2]) == merged["out"].encode("latin-1")
        assert stats.merged == merged["records_filtered"]
        assert stats.merged_written == (stats.merged if case["merge"].get("output", True) else 0)
        assert stats.ops["records_written"] + stats.merged_written == res["ops"]["records_written"]
        stats.ops["records_written"] += stats.merged_written
        stats.bp_out[0] += stats.bp_merged_written
    check_ops(case, stats, paired=True)
    assert bytes(outs[0]) == res["out1"].encode("latin-1")
    assert bytes(outs[1]) == res["out2"].encode("latin-1")
    assert stats.records == res["records"]
    assert list(stats.with_adapters) == [w or 0 for w in res["with_adapters"]]
    assert list(stats.bp_in) == res["bp_in"] and list(stats.bp_out) == res["bp_out"]
    assert stats.overflow == 0
    if res.get("corrected"):
        assert stats.records_corrected == res["corrected"]["records_corrected"]
        assert list(stats.bp_corrected) == res["corrected"]["bp_corrected"]
    for i, gold in enumerate(res["adapters"]):
        per_adapter = gold if isinstance(gold, list) else [gold]          # adapter mode: several adapters per read
        for a, g in enumerate(per_adapter):
            mine = stats.adapter_summary(i, a, WHERE[g["where"]])
            for key in ("lengths_front", "lengths_back", "errors_front", "errors_back", "adjacent_bases"):
                assert (key in g) == (key in mine), (case["label"], i, a, key)
                if key in g:
                    assert _str_keys(mine[key]) == g[key], (case["label"], i, a, key)
