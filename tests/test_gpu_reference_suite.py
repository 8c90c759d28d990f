"""The drop-in, dropped in: the reference's OWN test-suite (all of baseline/_ref/tests: test_align, test_adapters,
test_modifiers, the CLI goldens of test_atropos / test_paired, colorspace, filters, seqio ...) executed on the B200
with `atropos.align._align` served by atropos_b200 -- once through the per-call binding (module swap only) and once
through the batched binding (staged TrimPipeline.handle_records), see atropos_b200/integration.py.

Expected, both ways: what the unmodified reference gives here with its own Cython module -- 221 passed, 1 skipped
(the SRA test needs the network). The two tests that fork worker processes (`--threads N`) run in processes of their
own: a CUDA context does not survive fork(), so the parent must not have touched the GPU before it forks (with the
binding it never does: contexts are created on first use, in the worker). `IssueTests.test_issue68` is not collected
by the reference's pytest configuration (class name); it is run explicitly as an extra insert-aligner golden.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "run_reference_suite.py")
STAGED = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "atropos"))
FORKING = ["tests/test_paired.py::test_no_writer_process", "tests/test_paired.py::test_summary"]


def run_suite(mode, args, sim=False, timeout=1500):
    cmd = [sys.executable, RUNNER, "--mode", mode] + (["--sim"] if sim else []) + args
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    info = None
    for line in p.stdout.splitlines():
        if line.startswith("ATR_INTEGRATION "):
            info = json.loads(line[len("ATR_INTEGRATION "):])
    assert info is not None, p.stdout[-4000:]
    summary = [ln for ln in p.stdout.splitlines() if " passed" in ln or " failed" in ln or " error" in ln]
    return p.returncode, info, (summary[-1] if summary else ""), p.stdout


def counts(summary):
    out = {}
    for part in summary.replace("=", " ").split(","):
        bits = part.split()
        for i, b in enumerate(bits):
            if b in ("passed", "failed", "skipped", "error", "errors", "deselected") and i > 0 and bits[i - 1].isdigit():
                out[b] = int(bits[i - 1])
    return out


def check_whole_suite(mode, sim):
    deselect = []
    for t in FORKING:
        deselect += ["--deselect", t]
    rc, info, summary, text = run_suite(mode, ["tests", "-q"] + deselect, sim=sim)
    c = counts(summary)
    assert rc == 0 and c.get("failed", 0) == 0 and c.get("error", 0) == 0 and c.get("errors", 0) == 0, text[-6000:]
    assert c.get("passed") == 219 and c.get("skipped") == 1, summary
    assert info["shim"] and info["align_module"].startswith("atropos_b200")
    assert info["percall_locate"] > 100
    if mode == "batched":
        assert info["batched_batches"] > 50 and info["batched_reads"] > 800 and info["batched_gpu_calls"] > 50, info
    else:
        assert info["batched_batches"] == 0 and info["percall_locate"] > 1000, info
    if not sim:
        assert isinstance(info["gpu_launches"], int) and info["gpu_launches"] > 500, info
    # the two forking tests and the uncollected issue-68 golden, each in a fresh process
    for t in FORKING:
        rc, info, summary, text = run_suite(mode, [t, "-q"], sim=sim)
        assert rc == 0 and counts(summary).get("passed") == 1, text[-6000:]
    rc, info, summary, text = run_suite(mode, ["tests/test_paired.py", "-q", "-o", "python_classes=IssueTests", "-k",
                                               "test_issue68"], sim=sim)
    assert rc == 0 and counts(summary).get("passed") == 1, text[-6000:]
    if mode == "batched":
        assert info["batched_batches"] >= 1 and info["percall_batches"] == 0, info


@pytest.mark.gpu
@pytest.mark.skipif(not STAGED, reason="baseline/_ref not staged (python oracle/build_ref.py in the build container)")
@pytest.mark.parametrize("mode", ["percall", "batched"])
def test_reference_suite_on_gpu(mode):
    check_whole_suite(mode, sim=False)


@pytest.mark.skipif(not STAGED, reason="baseline/_ref not staged (python oracle/build_ref.py in the build container)")
@pytest.mark.parametrize("mode", ["percall", "batched"])
def test_reference_suite_glue_on_cpu_sim(mode):
    """Same tests, same binding code, device functions on the CPU (tests/simbackend.py): checks the Python glue here."""
    check_whole_suite(mode, sim=True)
