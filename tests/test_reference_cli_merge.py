"""The reference's own `trim` command line (unmodified, baseline/_ref) with --merge-overlapping --merged-output through
the batched binding of atropos_b200/integration.py: two GPU stages per batch (adapter stage, then ONE merge-alignment
call; the modifier's own code does the rest) and not a single per-call alignment. All three files must equal what the
reference wrote with its own Cython aligner (the golden cases of tests/golden/fastq_trim_pe.json.gz)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "run_reference_cli.py")
STAGED = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "atropos"))
CASES = ["merge_insert", "merge_insert_ragged_ops", "merge_correct_liberal", "merge_correct_n", "merge_adapter_mode_correct",
         "merge_discarded", "pair_filter_both_adapter_mode_merge"]


def run_case(case, sim):
    cmd = [sys.executable, RUNNER, "--case", case] + (["--sim"] if sim else [])
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    info = None
    for line in p.stdout.splitlines():
        if line.startswith("ATR_CLI "):
            info = json.loads(line[len("ATR_CLI "):])
    assert info is not None, p.stdout[-3000:]
    assert p.returncode == 0 and info["rc"] == 0 and all(info["same"].values()) and len(info["same"]) >= 2, info
    assert info["batched_batches"] >= 1 and info["batched_merge_batches"] >= 1 and info["percall_batches"] == 0, info
    assert info["percall_locate"] == 0, info          # MergeOverlapping's alignments all came from the batch call


@pytest.mark.skipif(not STAGED, reason="baseline/_ref not staged (python oracle/build_ref.py in the build container)")
@pytest.mark.parametrize("case", CASES[:3])
def test_cli_merge_glue_on_cpu_sim(case):
    run_case(case, sim=True)


@pytest.mark.gpu
@pytest.mark.skipif(not STAGED, reason="baseline/_ref not staged (python oracle/build_ref.py in the build container)")
@pytest.mark.parametrize("case", CASES)
def test_cli_merge_on_gpu(case):
    run_case(case, sim=False)
