"""Small single-end + paired-end FASTQ runs for ncu captures of the FASTQ kernels (tools only; bench.py measures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from atropos_b200 import synth, fastq
from atropos_b200.adapters import Adapter, BACK
from atropos_b200.align import InsertAligner
from atropos_b200.util import RandomMatchProbability
n, L = 2_000_000, 150
reads = synth.synth_se(n, L, seed=synth.seed_for(2), device="cpu").numpy()
text = synth.fastq_text(reads)
tr = fastq.FastqTrimmer([Adapter(synth.TRUSEQ_R1, BACK, max_error_rate=0.1, min_overlap=3)], times=1, max_len=L,
                        chunk_bytes=1 << 30, quality_cutoff=[20], trim_n=True, minimum_length=25)
out, st, c = tr.trim(text)
print("SE", st.records, st.with_adapters, st.ops)
if "--pe" in sys.argv:
    m = 1_000_000
    r1, r2 = synth.synth_pe(m, L, seed=synth.seed_for(3), device="cpu")
    rmp = RandomMatchProbability()
    kw = dict(max_error_rate=0.1, min_overlap=1, indel_cost=3, max_rmp=1e-6, match_probability=rmp)
    ptr = fastq.FastqPairTrimmer(Adapter(synth.TRUSEQ_R1, BACK, **kw), Adapter(synth.TRUSEQ_R2, BACK, **kw),
                                 InsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, match_probability=rmp,
                                               max_insert_mismatch_frac=0.1, max_adapter_mismatch_frac=0.1), max_len=L,
                                 chunk_bytes=1 << 30, trim_n=True, minimum_length=25)
    outs, pst, _ = ptr.trim(synth.fastq_text(r1.numpy()), synth.fastq_text(r2.numpy()))
    print("PE", pst.records, pst.insert_matches, pst.with_adapters, pst.ops)
