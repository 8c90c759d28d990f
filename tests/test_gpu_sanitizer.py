"""compute-sanitizer over every kernel family (tools/sanitizer_workload.py: K1 funnel both ways, Myers / anchored / general
kernels, a panel with rounds, K2, merge, packers, both FASTQ paths -- each also checked against the oracle / goldens).
memcheck and racecheck must report no errors. The summaries are what profiles/r2_sanitizer_*.txt record."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_compute_sanitizer_clean(tool):
    cs = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(cs):
        pytest.skip("compute-sanitizer not installed")
    p = subprocess.run([cs, "--tool", tool, "--error-exitcode", "86", sys.executable,
                        os.path.join(ROOT, "tools", "sanitizer_workload.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1700)
    tail = p.stdout[-3000:]
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "sanitizer_%s.txt" % tool), "w") as fh:
            fh.write(tail)
    assert "sanitizer workload ok" in p.stdout, tail
    assert p.returncode == 0, tail
    clean = "ERROR SUMMARY: 0 errors" if tool == "memcheck" else "RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings)"
    assert clean in p.stdout, tail
