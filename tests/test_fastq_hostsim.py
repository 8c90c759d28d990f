"""The FASTQ-in -> trimmed-FASTQ-out path (rows f-1/f-2/f-3) on the CPU: the device functions of
atropos_b200/csrc/fastq_core.cuh + the alignment funnel, run by tests/host_sim, against what the reference
command line produced (tests/golden/fastq_trim.json.gz): output text byte for byte, every statistic of the
report, and the reader's FormatError messages."""
import pytest

import fastq_cases
import hostsim
from atropos_b200 import fastq

CASES = fastq_cases.cases()


@pytest.mark.parametrize("case", CASES, ids=[c["label"] for c in CASES])
def test_against_reference_cli(case):
    adapters = fastq_cases.adapters_of(case)
    text = case["text"].encode("latin-1")
    res = case["result"]
    if "error" in res:
        with pytest.raises(fastq.FormatError) as ei:
            hostsim.trim_fastq(text, adapters, times=case["times"], **case.get("read_ops", {}))
        assert str(ei.value) == res["error"]
        return
    out, stats, consumed = hostsim.trim_fastq(text, adapters, times=case["times"], **case.get("read_ops", {}))
    assert consumed == len(text)
    fastq_cases.check(case, out, stats, adapters)


def test_streaming_chunks_reassemble():
    """final=False stops after the last complete record and reports `consumed`; feeding the rest gives the same
    output and statistics as one call"""
    case = [c for c in CASES if c["label"] == "panel_times3"][0]
    adapters = fastq_cases.adapters_of(case)
    text = case["text"].encode("latin-1")
    outs, stats, pos = [], None, 0
    for cut in (10_000, 33_333, 90_001, len(text)):
        final = cut == len(text)
        out, st, consumed = hostsim.trim_fastq(text[pos:cut], adapters, times=case["times"], final=final)
        outs.append(out)
        stats = st if stats is None else stats.merge(st)
        pos += consumed
    assert pos == len(text)
    fastq_cases.check(case, b"".join(outs), stats, adapters)


PE_CASES = fastq_cases.pe_cases()


@pytest.mark.parametrize("case", PE_CASES, ids=[c["label"] for c in PE_CASES])
def test_pe_against_reference_cli(case):
    a1, a2, ia = fastq_cases.pe_objects(case)
    t1, t2 = case["text1"].encode("latin-1"), case["text2"].encode("latin-1")
    res = case["result"]
    if "error" in res:
        with pytest.raises(fastq.FormatError) as ei:
            hostsim.trim_fastq_pe(t1, t2, a1, a2, ia, times=case.get("times", 1), mismatch_action=case.get("mismatch_action"), **fastq_cases.merge_kwargs(case), **case.get("read_ops", {}))
        assert str(ei.value) == res["error"]
        return
    outs, stats, consumed = hostsim.trim_fastq_pe(t1, t2, a1, a2, ia, times=case.get("times", 1), mismatch_action=case.get("mismatch_action"), **fastq_cases.merge_kwargs(case), **case.get("read_ops", {}))
    assert consumed == (len(t1), len(t2))
    fastq_cases.pe_check(case, outs, stats)
